set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_gputests_f.txt
tail -4 gpurun_out/r02_gputests_f.txt
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/r02_bench_f_default.json 2> gpurun_out/r02_bench_f.err
SES3D_TRI_EXACT=0 $B > gpurun_out/r02_bench_f_exact0.json 2>> gpurun_out/r02_bench_f.err
SES3D_TRI_EXACT=1 $B > gpurun_out/r02_bench_f_exact1.json 2>> gpurun_out/r02_bench_f.err
timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl_f.json > /dev/null 2>> gpurun_out/r02_bench_f.err
rm -f gpurun_out/*_chrome.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_f_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e_ms %.3f kms %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(f, "ERR", e)
d = json.load(open("gpurun_out/r02_tl_f.json"))
print("tl wall", d["wall_ms"], "span", d["gpu_span_ms"])
PY
tail -5 gpurun_out/r02_bench_f.err
