set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputests_b.txt
tail -4 gpurun_out/r02_gputests_b.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b_direct.json 2> gpurun_out/r02_bench_b.err
SES3D_RAGGED_DIRECT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_b_staged.json 2>> gpurun_out/r02_bench_b.err
for ch in 512 2048 4096; do SES3D_RAGGED_CHUNK=$ch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_b_direct_c$ch.json 2>> gpurun_out/r02_bench_b.err; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_b_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e %.3e e2e_ms %.3f lat %s kms %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("single_frame_call_p50_us"), d["roofline"]["kernel_ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_bench_b.err
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r02_topo.txt
