set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_gputests_i.txt
tail -3 gpurun_out/r02_gputests_i.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for W in 2 4 8 12 16; do
SES3D_TRI_WARPS=$W $B > gpurun_out/r02_bench_i_W${W}.json 2>> gpurun_out/r02_bench_i.err
done
SES3D_TRI_LOCKSTEP=0 SES3D_TRI_WARPS=2 $B > gpurun_out/r02_bench_i_L0W2.json 2>> gpurun_out/r02_bench_i.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_i_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e_ms %.3f kms %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_i.err
