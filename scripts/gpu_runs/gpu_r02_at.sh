cd $GRAFT_REPO_ROOT
for rep in 1 2; do
timeout 120 python scripts/extra_probe.py cfg3_fp64_lm 2>/dev/null | sed 's/^/new /'
SES3D_LIB=$GRAFT_REPO_ROOT/scripts/_variants/libses3d_old.so timeout 120 python scripts/extra_probe.py cfg3_fp64_lm 2>/dev/null | sed 's/^/old /'
done
