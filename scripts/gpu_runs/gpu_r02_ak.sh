cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r02_split_$tag.json 2>> gpurun_out/r02_split.err; }
rm -f gpurun_out/r02_split_*.json
for rep in 1 2; do
run s1_$rep SES3D_DEVICE_SPLIT=1
run s2_$rep SES3D_DEVICE_SPLIT=2
run s3_$rep SES3D_DEVICE_SPLIT=3
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_split_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('r02_split_')[1], "dev ms %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"])
PY
