cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gputests_r.txt
tail -3 gpurun_out/r02_gputests_r.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/r02_bench_r_default.json 2>> gpurun_out/r02_bench_r.err
for T in 1 0; do
SES3D_PAIRS_TILED=$T $B --workload cfg4_crowd64x20 --frames 512 > gpurun_out/r02_bench_r_cfg4_T$T.json 2>> gpurun_out/r02_bench_r.err
SES3D_PAIRS_TILED=$T $B --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_r_dense_T$T.json 2>> gpurun_out/r02_bench_r.err
SES3D_PAIRS_TILED=$T $B --workload cfg5_ring8x4 > gpurun_out/r02_bench_r_cfg5_T$T.json 2>> gpurun_out/r02_bench_r.err
done
SES3D_PAIRS_TILED=0 $B > gpurun_out/r02_bench_r_default_T0.json 2>> gpurun_out/r02_bench_r.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_cfg4b.csv python scripts/profile_step.py --workload cfg4_crowd64x20 --frames 512 --steps 2 > gpurun_out/r02_b_launch_cfg4b.log 2>&1
python - <<'PY'
import json, glob, csv
for f in sorted(glob.glob("gpurun_out/r02_bench_r_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
rows = list(csv.reader(open("gpurun_out/r02_launches_cfg4b.csv")))
hdr = None
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        dd = dict(zip(hdr, r))
        if "ses3d" in dd["Kernel Name"]: print(dd["ID"], dd["Kernel Name"][:40], dd["Grid Size"], dd["Block Size"], dd["Metric Value"])
PY
tail -5 gpurun_out/r02_bench_r.err
