cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_frame_graph.py -q -m gpu -x 2>&1 | tail -3
for rep in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_ar_$rep.json 2> gpurun_out/r02_bench_ar.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_ar_$rep.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --workload dense_ring16x6 --frames 4096 > gpurun_out/r02_bench_ar_dense.json 2>> gpurun_out/r02_bench_ar.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_ar_dense.json').read().strip().splitlines()[-1]); print('dense ms', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
