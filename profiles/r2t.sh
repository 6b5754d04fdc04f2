#!/bin/bash
# round 2, GPU call T: source-level ncu captures of k_forward_wp and k_backward (config 4, one working launch each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in k_forward_wp k_backward; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2t_$k -f python profiles/prof_c4.py > gpurun_out/r2t_$k.log 2>&1
  ncu -i gpurun_out/r2t_$k.ncu-rep --page raw --csv > gpurun_out/r2t_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2t_$k.ncu-rep --page source --csv > gpurun_out/r2t_${k}_source.csv 2>/dev/null
  rm -f gpurun_out/r2t_$k.ncu-rep
  tail -n 2 gpurun_out/r2t_$k.log
done
ls -la gpurun_out/ | grep r2t
