cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r02_e2e_$tag.json 2>> gpurun_out/r02_e2e.err; }
for rep in 1 2; do
run base_$rep A=1
run oneup_$rep SES3D_RAGGED_ONE_UP=1
run onedown_$rep SES3D_RAGGED_ONE_DOWN=1
run both_$rep SES3D_RAGGED_ONE_UP=1 SES3D_RAGGED_ONE_DOWN=1
run slots3_$rep SES3D_RAGGED_SLOTS=3
run slots4_$rep SES3D_RAGGED_SLOTS=4
run both_c768_$rep SES3D_RAGGED_ONE_UP=1 SES3D_RAGGED_ONE_DOWN=1 SES3D_RAGGED_CHUNK=768
run both_c1536_$rep SES3D_RAGGED_ONE_UP=1 SES3D_RAGGED_ONE_DOWN=1 SES3D_RAGGED_CHUNK=1536
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_e2e_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('r02_e2e_')[1], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "dev ms %.3f" % d["ms_per_step"])
PY
