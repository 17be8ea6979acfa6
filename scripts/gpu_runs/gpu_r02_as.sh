cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for rep in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_as_$rep.json 2> gpurun_out/r02_bench_as.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_as_$rep.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])"
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --workload cfg4_crowd64x20 --frames 512 > gpurun_out/r02_bench_as_crowd.json 2>> gpurun_out/r02_bench_as.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_as_crowd.json').read().strip().splitlines()[-1]); print('crowd ms', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
timeout 300 python scripts/latency_kernels.py 2>&1 | grep -E "k_rounds|k_triangulate|device span"
