set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_gputests_e.txt
tail -4 gpurun_out/r02_gputests_e.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_e_full.json 2> gpurun_out/r02_bench_e.err
SES3D_RAGGED_DIRECT=0 timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl_staged.json > /dev/null 2>> gpurun_out/r02_bench_e.err
SES3D_RAGGED_DIRECT=1 timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl_direct.json > /dev/null 2>> gpurun_out/r02_bench_e.err
SES3D_RAGGED_DIRECT=0 SES3D_TRI_EXACT=0 timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl_staged_exact0.json > /dev/null 2>> gpurun_out/r02_bench_e.err
SES3D_RAGGED_DIRECT=0 SES3D_RAGGED_CHUNK=2048 timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl_staged_c2048.json > /dev/null 2>> gpurun_out/r02_bench_e.err
SES3D_RAGGED_DIRECT=0 SES3D_RAGGED_CHUNK=4096 timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl_staged_c4096.json > /dev/null 2>> gpurun_out/r02_bench_e.err
rm -f gpurun_out/*_chrome.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_tl_*.json")):
    d = json.load(open(f))
    print(f, "wall", [round(x,2) for x in d["wall_ms"]], "span", round(d["gpu_span_ms"],2), "kern_union", round(d["all_kernels_busy_union_ms"],2))
    for k, v in d["rows"].items():
        print("   %-28s n=%4d sum=%7.2f union=%7.2f  [%6.2f .. %6.2f]" % (k[:28], v["n"], v["sum_ms"], v["busy_union_ms"], v["first_start_ms"], v["last_end_ms"]))
d = json.loads(open("gpurun_out/r02_bench_e_full.json").read().strip().splitlines()[-1])
print("value %.3e ms %.3f e2e_ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d["roofline"]["kernel_ms_per_step"], "frac", d["roofline"]["frac"])
for k, v in d.get("extra", {}).items():
    print(k, json.dumps(v)[:600])
PY
tail -5 gpurun_out/r02_bench_e.err
