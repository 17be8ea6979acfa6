set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputests_a.txt
timeout 1500 python scripts/soak_parity.py --scale 1.0 > gpurun_out/r02_soak_a.jsonl 2> gpurun_out/r02_soak_a.err
tail -3 gpurun_out/r02_gputests_a.txt
cat gpurun_out/r02_soak_a.jsonl | cut -c1-900
