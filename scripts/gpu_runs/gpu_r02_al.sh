cd $GRAFT_REPO_ROOT
timeout 200 python scripts/pcie_probe_concurrent.py 2>/dev/null | tail -1
B="timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r02_e2e2_$tag.json 2>> gpurun_out/r02_e2e2.err; }
rm -f gpurun_out/r02_e2e2_*.json
for rep in 1 2; do
run new_$rep A=1
run old_$rep SES3D_RAGGED_ONE_UP=0 SES3D_RAGGED_ONE_DOWN=0 SES3D_RAGGED_CHUNK=1024
run new_c1024_$rep SES3D_RAGGED_CHUNK=1024
run onedown_only_$rep SES3D_RAGGED_ONE_UP=0
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_e2e2_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('r02_e2e2_')[1], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "dev ms %.3f" % d["ms_per_step"])
PY
