cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_frame_graph.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python scripts/latency_kernels.py 2>&1 | grep -v Warn | tee gpurun_out/latency_kernels_r02.txt
timeout 300 python scripts/latency_probe.py 2>&1 | tee gpurun_out/latency_r02b.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 30 -c 2 -o gpurun_out/r02_rounds_1frame -f python scripts/latency_one.py > gpurun_out/ncu_rounds1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_triangulate -s 30 -c 2 -o gpurun_out/r02_tri_1frame -f python scripts/latency_one.py > gpurun_out/ncu_tri1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
