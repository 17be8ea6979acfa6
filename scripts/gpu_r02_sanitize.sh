cd $GRAFT_REPO_ROOT
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_smoke.py > gpurun_out/r02_san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke done" gpurun_out/r02_san_$tool.log | tail -3
done
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_prior.py > gpurun_out/r02_san_prior_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY" gpurun_out/r02_san_prior_racecheck.log | tail -1
