cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prior.py -q -m gpu -k "ragged or chunk or multi" 2>&1 | tail -3
B="timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra"
rm -f gpurun_out/r02_e2e3_*.json
for rep in 1 2 3; do
$B > gpurun_out/r02_e2e3_new_$rep.json 2>> gpurun_out/r02_e2e3.err
SES3D_LIB=$GRAFT_REPO_ROOT/scripts/_variants/libses3d_prev.so $B > gpurun_out/r02_e2e3_prev_$rep.json 2>> gpurun_out/r02_e2e3.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_e2e3_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('r02_e2e3_')[1], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "dev ms %.3f" % d["ms_per_step"])
PY
