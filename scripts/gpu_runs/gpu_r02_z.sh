cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for rep in 1 2; do for v in M0 M1; do
SES3D_LIB=$GRAFT_REPO_ROOT/scripts/_variants/libses3d_$v.so $B > gpurun_out/r02_bench_z_${v}_$rep.json 2>> gpurun_out/r02_bench_z.err
done; done
SES3D_LIB=$GRAFT_REPO_ROOT/scripts/_variants/libses3d_M1.so $B --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_z_dense_M1.json 2>> gpurun_out/r02_bench_z.err
SES3D_LIB=$GRAFT_REPO_ROOT/scripts/_variants/libses3d_M0.so $B --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_z_dense_M0.json 2>> gpurun_out/r02_bench_z.err
SES3D_LIB=$GRAFT_REPO_ROOT/scripts/_variants/libses3d_M1.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_c_abi_example.py -m gpu -q 2>&1 | tail -2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_z_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
PY
