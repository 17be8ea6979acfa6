cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gputests_t.txt
tail -3 gpurun_out/r02_gputests_t.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for S in 0 1 2 4 8 16; do
SES3D_PAIRS_SPLIT=$S $B --workload cfg4_crowd64x20 --frames 512 > gpurun_out/r02_bench_t_cfg4_S$S.json 2>> gpurun_out/r02_bench_t.err
done
for A in 128 256; do SES3D_PAIRS_SPLIT=4 SES3D_ASSOC_THREADS=$A $B --workload cfg4_crowd64x20 --frames 512 > gpurun_out/r02_bench_t_cfg4_S4_A$A.json 2>> gpurun_out/r02_bench_t.err; done
python - <<'PY'
import json, glob, csv
for f in sorted(glob.glob("gpurun_out/r02_bench_t_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_t.err
