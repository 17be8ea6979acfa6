cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_gputests_n.txt
tail -4 gpurun_out/r02_gputests_n.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for T in 0 128 256; do
SES3D_REPROJ_THREADS=$T $B > gpurun_out/r02_bench_n_T${T}.json 2>> gpurun_out/r02_bench_n.err
done
SES3D_REPROJ_THREADS=128 SES3D_ROUNDS_LOCKSTEP=0 $B > gpurun_out/r02_bench_n_T128_RL0.json 2>> gpurun_out/r02_bench_n.err
SES3D_REPROJ_THREADS=128 SES3D_ROUNDS_WARPS=8 $B > gpurun_out/r02_bench_n_T128_RW8.json 2>> gpurun_out/r02_bench_n.err
SES3D_REPROJ_THREADS=128 SES3D_ROUNDS_WARPS=2 $B > gpurun_out/r02_bench_n_T128_RW2.json 2>> gpurun_out/r02_bench_n.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_n_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_n.err
