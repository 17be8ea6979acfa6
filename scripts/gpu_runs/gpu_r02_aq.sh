cd $GRAFT_REPO_ROOT
for cfg in "A=1" "SES3D_TRI_LOCKSTEP=0" "SES3D_TRI_LOCKSTEP=0 SES3D_TRI_WARPS=2" "SES3D_TRI_WARPS=2" "SES3D_TRI_LOCKSTEP=0 SES3D_TRI_WARPS=8"; do
echo "== $cfg"
env $cfg timeout 300 python scripts/latency_kernels.py 2>&1 | grep -E "k_triangulate|device span"
done
