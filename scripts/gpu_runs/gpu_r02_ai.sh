cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_frame_graph.py -q -m gpu 2>&1 | tail -2
run() { tag=$1; shift; env "$@" timeout 300 python scripts/bench_prior.py --no-cpu --sequences $SEQ --steps 6 > gpurun_out/r02_prior_$tag.json 2>> gpurun_out/r02_prior.err; }
for SEQ in 1 64 256 512; do
run s${SEQ}_default A=1
run s${SEQ}_g1w6 SES3D_PRIOR_GROUP=1 SES3D_PRIOR_WARPS=6
run s${SEQ}_g2w3 SES3D_PRIOR_GROUP=2 SES3D_PRIOR_WARPS=3
run s${SEQ}_g1w4 SES3D_PRIOR_GROUP=1 SES3D_PRIOR_WARPS=4
run s${SEQ}_g3w4 SES3D_PRIOR_GROUP=3 SES3D_PRIOR_WARPS=4
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_prior_s*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r02_prior_')[1], "ms %.3f" % d["ms_per_step"], "single-msg us %.1f" % d["single_message_call_p50_us"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r02_prior.err
