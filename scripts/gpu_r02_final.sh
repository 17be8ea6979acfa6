# Round-2 record run: tests, smoke, both bench arms, launch list, ncu --set full of the hot kernels, soak, latency.
cd $GRAFT_REPO_ROOT
T=${1:-r02}
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/${T}_gpu_tests.log; cat gpurun_out/${T}_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_${T}_reference.json 2> gpurun_out/${T}_bench_ref_err.log
timeout 1200 python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/${T}_bench_err.log; tail -2 gpurun_out/${T}_bench_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python scripts/profile_step.py --steps 6 > gpurun_out/${T}_b_launch.log 2>&1
for k in k_triangulate k_finproj k_pairs k_rounds; do
  SES3D_DEVICE_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_${T} python scripts/profile_step.py --steps 4 > gpurun_out/${T}_b_$k.log 2>&1
  ncu -i gpurun_out/prof_${k}_${T}.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${k}_${T}.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${T}_${k}_src.csv 2>/dev/null
done
rm -f gpurun_out/prof_k_pairs_${T}.ncu-rep gpurun_out/prof_k_rounds_${T}.ncu-rep
timeout 300 python scripts/e2e_timeline.py --out gpurun_out/${T}_tl.json > /dev/null 2>> gpurun_out/${T}_bench_err.log
rm -f gpurun_out/*_chrome.json
timeout 300 python scripts/latency_probe.py > gpurun_out/latency_${T}.txt 2>&1; tail -3 gpurun_out/latency_${T}.txt
timeout 300 python scripts/latency_kernels.py 2>&1 | grep -v -i warn > gpurun_out/${T}_latency_kernels.txt; tail -2 gpurun_out/${T}_latency_kernels.txt
timeout 200 python scripts/pcie_probe_concurrent.py > gpurun_out/${T}_link_n1.json 2>/dev/null; timeout 200 python scripts/pcie_probe_pieces.py 2>&1 | grep -v -i warn > gpurun_out/${T}_link_pieces.txt
python - <<PY
import json
for f in ['bench_${T}_reference','bench_${T}']:
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d.get('frames_per_sec'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'))
    for k,v in d.get('extra',{}).items(): print('  ',k, {kk: v.get(kk) for kk in ('ms_per_step','frames_per_sec','value')}, v.get('parity',{}).get('ok'), v.get('roofline',{}).get('kernel'), v.get('roofline',{}).get('frac'), v.get('error'))
PY
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_smoke.py > gpurun_out/${T}_san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke done" gpurun_out/${T}_san_$tool.log | tail -3
done
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_prior.py > gpurun_out/${T}_san_prior_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY" gpurun_out/${T}_san_prior_racecheck.log | tail -1
timeout 900 python scripts/bench_prior.py > gpurun_out/bench_prior_${T}.json 2> gpurun_out/${T}_bench_prior_err.log; python -c "
import json; d=json.loads(open('gpurun_out/bench_prior_${T}.json').read().strip().splitlines()[-1]); print('prior', d['value'], d.get('ms_per_step'), d.get('e2e',{}).get('value'))"
timeout 900 python scripts/bench_chain.py > gpurun_out/bench_chain_${T}.json 2> gpurun_out/${T}_err_chain.log; python -c "
import json; d=json.loads(open('gpurun_out/bench_chain_${T}.json').read().strip().splitlines()[-1]); print('chain', d.get('value'), d.get('frames_per_sec'), d.get('ms_per_step'))"
# the large-sample parity soak last (the longest item; everything above is already on disk if the budget runs out)
timeout 420 python scripts/soak_parity.py > gpurun_out/${T}_soak.jsonl 2> gpurun_out/${T}_soak.err; cut -c1-300 gpurun_out/${T}_soak.jsonl
