cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_gputests_o.txt
tail -4 gpurun_out/r02_gputests_o.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/r02_bench_o_default.json 2>> gpurun_out/r02_bench_o.err
for ch in 512 768 1536; do SES3D_RAGGED_CHUNK=$ch $B > gpurun_out/r02_bench_o_c$ch.json 2>> gpurun_out/r02_bench_o.err; done
timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl_o.json > /dev/null 2>> gpurun_out/r02_bench_o.err
rm -f gpurun_out/*_chrome.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_o_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
d = json.load(open("gpurun_out/r02_tl_o.json"))
print("tl wall", d["wall_ms"], "span", d["gpu_span_ms"])
for k, v in d["rows"].items():
    print("   %-28s n=%4d sum=%7.2f union=%7.2f  [%6.2f .. %6.2f]" % (k[:28], v["n"], v["sum_ms"], v["busy_union_ms"], v["first_start_ms"], v["last_end_ms"]))
PY
tail -5 gpurun_out/r02_bench_o.err
