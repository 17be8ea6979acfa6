cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for rep in 1 2; do for v in J1P1U0 J0P1U0 J0P0U0 J0P0U1 J1P1U1 J1P0U1; do
SES3D_LIB=$GRAFT_REPO_ROOT/scripts/_variants/libses3d_$v.so $B > gpurun_out/r02_bench_l_${v}_$rep.json 2>> gpurun_out/r02_bench_l.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_l_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms %.3f tri %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"]["triangulate"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_l.err
