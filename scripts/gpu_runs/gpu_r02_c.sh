set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_gputests_c.txt
tail -6 gpurun_out/r02_gputests_c.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c_direct.json 2> gpurun_out/r02_bench_c.err
SES3D_RAGGED_DIRECT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_c_staged.json 2>> gpurun_out/r02_bench_c.err
for ch in 2048 4096; do SES3D_RAGGED_CHUNK=$ch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_c_direct_c$ch.json 2>> gpurun_out/r02_bench_c.err; done
SES3D_DEVICE_SPLIT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_c_split1.json 2>> gpurun_out/r02_bench_c.err
SES3D_DEVICE_SPLIT=3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_c_split3.json 2>> gpurun_out/r02_bench_c.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_c_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e %.3e e2e_ms %.3f lat %s kms %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("single_frame_call_p50_us"), {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_c.err
python scripts/pcie_probe.py > gpurun_out/r02_pcie_probe.txt 2>&1; tail -12 gpurun_out/r02_pcie_probe.txt
