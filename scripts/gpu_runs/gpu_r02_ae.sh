cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
SES3D_PRIOR_SMALL_MSGS=0 timeout 600 python scripts/bench_prior.py > gpurun_out/bench_prior_small0.json 2> gpurun_out/bench_prior_small0.err
timeout 600 python scripts/bench_prior.py > gpurun_out/bench_prior_r02.json 2> gpurun_out/bench_prior_r02.err
python - <<'PY'
import json
for f in ['bench_prior_small0','bench_prior_r02']:
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['single_message_call_p50_us'])
PY
