cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_gputests_w.txt; tail -3 gpurun_out/r02_gputests_w.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for i in 1 2; do $B > gpurun_out/r02_bench_w_$i.json 2>> gpurun_out/r02_bench_w.err; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_w_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
PY
timeout 1500 compute-sanitizer --tool synccheck python scripts/sanitize_smoke.py > gpurun_out/r02_san_synccheck.log 2>&1
grep -E "ERROR SUMMARY|sanitize smoke done" gpurun_out/r02_san_synccheck.log | tail -2
