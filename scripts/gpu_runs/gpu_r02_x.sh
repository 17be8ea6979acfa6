cd $GRAFT_REPO_ROOT
for k in k_triangulate k_finproj; do
  SES3D_DEVICE_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_r02b python scripts/profile_step.py --steps 4 > gpurun_out/r02b_b_$k.log 2>&1
  ncu -i gpurun_out/prof_${k}_r02b.ncu-rep --page raw --csv > gpurun_out/r02b_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${k}_r02b.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02b_${k}_src.csv 2>/dev/null
  rm -f gpurun_out/prof_${k}_r02b.ncu-rep
done
