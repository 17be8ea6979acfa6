cd $GRAFT_REPO_ROOT
timeout 120 python scripts/extra_probe.py cfg3_fp64_lm 2>/dev/null | sed 's/^/new /'
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp64 or parameter_variants or lm_refinement" 2>&1 | tail -2
timeout 120 python scripts/extra_probe.py cfg3_fp64_lm 2>/dev/null | sed 's/^/new /'
