cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_gputests_u.txt; tail -3 gpurun_out/r02_gputests_u.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/r02_bench_u_default.json 2>> gpurun_out/r02_bench_u.err
$B --workload cfg4_crowd64x20 --frames 512 > gpurun_out/r02_bench_u_cfg4.json 2>> gpurun_out/r02_bench_u.err
$B --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_u_dense.json 2>> gpurun_out/r02_bench_u.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_u_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
PY
bash scripts/gpu_r02_sanitize.sh
