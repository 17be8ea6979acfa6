set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python scripts/profile_step.py --steps 6 > gpurun_out/r02_b_launch.log 2>&1
for k in k_triangulate k_finproj k_pairs k_rounds; do
  SES3D_DEVICE_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_prof_$k python scripts/profile_step.py --steps 4 > gpurun_out/r02_b_$k.log 2>&1
  ncu -i gpurun_out/r02_prof_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_prof_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02_${k}_src.csv 2>/dev/null
  ls -la gpurun_out/r02_prof_$k.ncu-rep
done
rm -f gpurun_out/r02_prof_k_pairs.ncu-rep gpurun_out/r02_prof_k_rounds.ncu-rep
du -sh gpurun_out
