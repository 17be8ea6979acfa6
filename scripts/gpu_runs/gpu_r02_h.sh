set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_gputests_h.txt
tail -3 gpurun_out/r02_gputests_h.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for L in 1 0; do for W in 2 4 8; do
SES3D_TRI_LOCKSTEP=$L SES3D_TRI_WARPS=$W $B > gpurun_out/r02_bench_h_L${L}_W${W}.json 2>> gpurun_out/r02_bench_h.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_h_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e_ms %.3f kms %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_h.err
