cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_gputests_p.txt
tail -4 gpurun_out/r02_gputests_p.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/r02_bench_p_default.json 2>> gpurun_out/r02_bench_p.err
for ch in 768 1536 2048; do SES3D_RAGGED_CHUNK=$ch $B > gpurun_out/r02_bench_p_c$ch.json 2>> gpurun_out/r02_bench_p.err; done
timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl_p.json > /dev/null 2>> gpurun_out/r02_bench_p.err
rm -f gpurun_out/*_chrome.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_cfg4.csv python scripts/profile_step.py --workload cfg4_crowd64x20 --frames 512 --steps 3 > gpurun_out/r02_b_launch_cfg4.log 2>&1
python - <<'PY'
import json, glob, csv
for f in sorted(glob.glob("gpurun_out/r02_bench_p_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
d = json.load(open("gpurun_out/r02_tl_p.json"))
print("tl wall", d["wall_ms"], "span", d["gpu_span_ms"])
for k, v in d["rows"].items():
    print("   %-28s n=%4d sum=%7.2f union=%7.2f  [%6.2f .. %6.2f]" % (k[:28], v["n"], v["sum_ms"], v["busy_union_ms"], v["first_start_ms"], v["last_end_ms"]))
rows = list(csv.reader(open("gpurun_out/r02_launches_cfg4.csv")))
hdr = None
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        dd = dict(zip(hdr, r)); print(dd["ID"], dd["Kernel Name"][:40], dd["Grid Size"], dd["Block Size"], dd["Metric Value"])
PY
tail -5 gpurun_out/r02_bench_p.err
