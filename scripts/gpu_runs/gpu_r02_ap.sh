cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_frame_graph.py tests/test_ros_shim.py -q -m gpu 2>&1 | tail -3
for tool in initcheck memcheck; do
timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_smoke.py > gpurun_out/r02_san_$tool.log 2>&1
echo "== $tool"; grep -E "ERROR SUMMARY|sanitize smoke done" gpurun_out/r02_san_$tool.log | tail -3
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_ap.json 2> gpurun_out/r02_bench_ap.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_ap.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
