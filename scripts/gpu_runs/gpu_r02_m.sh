cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_gputests_m.txt
tail -4 gpurun_out/r02_gputests_m.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for T in 0 128 256 512; do for S in 6 3; do
SES3D_REPROJ_THREADS=$T SES3D_REPROJ_SCAP=$S $B > gpurun_out/r02_bench_m_T${T}_S${S}.json 2>> gpurun_out/r02_bench_m.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_m_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_m.err
