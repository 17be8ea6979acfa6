cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_gputests_v.txt; tail -3 gpurun_out/r02_gputests_v.txt
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for W in 4 8 2; do SES3D_TRI_WARPS=$W $B > gpurun_out/r02_bench_v_W$W.json 2>> gpurun_out/r02_bench_v.err; done
SES3D_TRI_LOCKSTEP=0 SES3D_TRI_WARPS=2 $B > gpurun_out/r02_bench_v_L0W2.json 2>> gpurun_out/r02_bench_v.err
SES3D_TRI_WARPS=8 $B --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_v_dense_W8.json 2>> gpurun_out/r02_bench_v.err
SES3D_TRI_WARPS=4 $B --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_v_dense_W4.json 2>> gpurun_out/r02_bench_v.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_v_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
PY
for tool in synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_smoke.py > gpurun_out/r02_san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke done" gpurun_out/r02_san_$tool.log | tail -3
done
