set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_gputests_d.txt
tail -8 gpurun_out/r02_gputests_d.txt
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/r02_bench_d_default.json 2> gpurun_out/r02_bench_d.err
SES3D_TRI_EXACT=0 $B > gpurun_out/r02_bench_d_exact0.json 2>> gpurun_out/r02_bench_d.err
SES3D_TRI_EXACT=1 $B > gpurun_out/r02_bench_d_exact1.json 2>> gpurun_out/r02_bench_d.err
SES3D_TRI_EXACT=0 SES3D_TRI_DYNAMIC=0 $B > gpurun_out/r02_bench_d_exact0_static.json 2>> gpurun_out/r02_bench_d.err
SES3D_TRI_DYNAMIC=0 $B > gpurun_out/r02_bench_d_static.json 2>> gpurun_out/r02_bench_d.err
SES3D_ASSOC_THREADS=96 $B > gpurun_out/r02_bench_d_at96.json 2>> gpurun_out/r02_bench_d.err
SES3D_ASSOC_THREADS=256 $B > gpurun_out/r02_bench_d_at256.json 2>> gpurun_out/r02_bench_d.err
SES3D_ROUNDS_WARPS=2 $B > gpurun_out/r02_bench_d_rw2.json 2>> gpurun_out/r02_bench_d.err
SES3D_RAGGED_DIRECT=0 $B > gpurun_out/r02_bench_d_staged.json 2>> gpurun_out/r02_bench_d.err
$B --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_d_dense.json 2>> gpurun_out/r02_bench_d.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_d_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e_ms %.3f kms %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_d.err
