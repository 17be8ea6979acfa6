set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_gputests_k.txt
tail -12 gpurun_out/r02_gputests_k.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for W in 4 2; do
SES3D_TRI_WARPS=$W $B > gpurun_out/r02_bench_k_W${W}.json 2>> gpurun_out/r02_bench_k.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_k_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e_ms %.3f kms %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_k.err
