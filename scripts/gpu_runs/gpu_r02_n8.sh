cd $GRAFT_REPO_ROOT
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR scripts/pcie_probe_concurrent.py > gpurun_out/r02_link_n$N.json 2> gpurun_out/r02_link_n$N.err; tail -1 gpurun_out/r02_link_n$N.json
timeout 300 python scripts/pcie_probe_concurrent.py > gpurun_out/r02_link_n1.json 2>> gpurun_out/r02_link_n$N.err; tail -1 gpurun_out/r02_link_n1.json
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err; tail -c 1500 gpurun_out/r02_scale_n$N.json
