set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/gpu_tests_r01_final.log; cat gpurun_out/gpu_tests_r01_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01.log 2>&1; tail -3 gpurun_out/smoke_r01.log
python bench.py --impl reference > gpurun_out/bench_r01_reference.json 2>gpurun_out/bench_ref_err.log
python bench.py > gpurun_out/bench_r01.json 2>gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python bench.py --workload dense_ring16x6 --frames 4096 > gpurun_out/bench_r01_dense.json 2>>gpurun_out/bench_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py --steps 6 > gpurun_out/b_launch.log 2>&1
for k in k_associate k_triangulate k_finalize k_reproject; do
  SES3D_DEVICE_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_r01 python scripts/profile_step.py --steps 4 > gpurun_out/b_$k.log 2>&1
done
python -c "
import json
for f in ['bench_r01_reference','bench_r01','bench_r01_dense']:
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d.get('frames_per_sec'), d.get('ms_per_step'), d.get('e2e',{}).get('value'))
"
python scripts/run_configs.py --out gpurun_out/configs_r01.json > gpurun_out/configs_log.txt 2>&1; tail -2 gpurun_out/configs_log.txt
python scripts/bench_prior.py > gpurun_out/bench_prior_r01.json 2>gpurun_out/bench_prior_err.log
python scripts/bench_chain.py > gpurun_out/bench_chain_r01.json 2>gpurun_out/err_chain.log
python scripts/latency_probe.py > gpurun_out/latency_r01.txt 2>&1; tail -3 gpurun_out/latency_r01.txt
