cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for S in 1 2 3; do SES3D_DEVICE_SPLIT=$S $B > gpurun_out/r02_bench_y_split$S.json 2>> gpurun_out/r02_bench_y.err; done
for W in 2 4; do SES3D_ROUNDS_WARPS=$W $B > gpurun_out/r02_bench_y_rw$W.json 2>> gpurun_out/r02_bench_y.err; done
for S in 3 4 8; do SES3D_REPROJ_SCAP=$S $B > gpurun_out/r02_bench_y_scap$S.json 2>> gpurun_out/r02_bench_y.err; done
for A in 96 192 256; do SES3D_ASSOC_THREADS=$A $B > gpurun_out/r02_bench_y_at$A.json 2>> gpurun_out/r02_bench_y.err; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_y_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
PY
tail -3 gpurun_out/r02_bench_y.err
