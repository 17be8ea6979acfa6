cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python scripts/latency_kernels.py 2>&1 | grep -v -i warn | tee gpurun_out/latency_kernels_r02.txt
timeout 300 python scripts/latency_probe.py 2>&1 | tee gpurun_out/latency_r02b.txt
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_smoke.py > gpurun_out/r02_san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke done" gpurun_out/r02_san_$tool.log | tail -3
done
