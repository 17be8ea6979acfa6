cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_prior.py tests/test_demo_chain.py tests/test_ros_shim.py tests/test_c_abi_example.py -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/bench_prior.py --no-cpu --steps 6 > gpurun_out/r02_prior_final.json 2>> gpurun_out/r02_prior.err
timeout 300 python scripts/bench_chain.py > gpurun_out/bench_chain_r02.json 2>> gpurun_out/r02_prior.err
timeout 300 python scripts/bench_chain.py --streams 128 > gpurun_out/bench_chain_r02_s128.json 2>> gpurun_out/r02_prior.err
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_prior.py 2>&1 | grep -E "RACECHECK SUMMARY|done" | tail -2
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_prior.py 2>&1 | grep -E "ERROR SUMMARY|done" | tail -2
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_prior_final.json").read().strip().splitlines()[-1])
print("prior ms %.3f single-msg us %.1f" % (d["ms_per_step"], d["single_message_call_p50_us"]))
for f in ("bench_chain_r02", "bench_chain_r02_s128"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d[k] for k in d if k in ("value", "ms_per_step", "frames_per_sec", "stage_ms")})
PY
