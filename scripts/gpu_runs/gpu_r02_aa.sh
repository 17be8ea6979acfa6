cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for T in 1 0; do
SES3D_PAIRS_BLOCKED=$T $B > gpurun_out/r02_bench_aa_cfg2_B$T.json 2>> gpurun_out/r02_bench_aa.err
SES3D_PAIRS_BLOCKED=$T $B --workload cfg4_crowd64x20 --frames 512 > gpurun_out/r02_bench_aa_cfg4_B$T.json 2>> gpurun_out/r02_bench_aa.err
SES3D_PAIRS_BLOCKED=$T $B --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_aa_dense_B$T.json 2>> gpurun_out/r02_bench_aa.err
SES3D_PAIRS_BLOCKED=$T $B --workload cfg5_ring8x4 > gpurun_out/r02_bench_aa_cfg5_B$T.json 2>> gpurun_out/r02_bench_aa.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_aa_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
PY
tail -3 gpurun_out/r02_bench_aa.err
