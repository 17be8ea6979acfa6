cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/bench_r02_reference.json 2> gpurun_out/r02_bench_ref_err.log
timeout 1200 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/r02_bench_err.log; tail -2 gpurun_out/r02_bench_err.log
python - <<'PY'
import json
for f in ['bench_r02_reference','bench_r02']:
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d.get('frames_per_sec'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('e2e',{}).get('ms_per_step'), d.get('roofline',{}).get('frac'))
    for k,v in d.get('extra',{}).items(): print('  ',k, {kk: v.get(kk) for kk in ('ms_per_step','frames_per_sec','value')}, v.get('parity',{}).get('ok'), v.get('roofline',{}).get('kernel'), v.get('roofline',{}).get('frac'), v.get('error'))
PY
