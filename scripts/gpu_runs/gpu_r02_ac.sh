cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_frame_graph.py tests/test_gpu_parity.py tests/test_ros_shim.py -q -m gpu -x 2>&1 | tail -15
timeout 300 python scripts/latency_kernels.py 2>&1 | tee gpurun_out/latency_kernels_r02.txt
SES3D_ROUNDS_BLOCK=1 timeout 300 python scripts/latency_kernels.py 2>&1 | tee gpurun_out/latency_kernels_r02_block.txt
