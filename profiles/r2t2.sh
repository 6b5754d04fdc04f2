#!/bin/bash
# round 2: source-level ncu capture of k_forward_wp on the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
k=k_forward_wp
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2t2_$k -f python profiles/prof_c4.py > gpurun_out/r2t2_$k.log 2>&1
ncu -i gpurun_out/r2t2_$k.ncu-rep --page source --csv > gpurun_out/r2t2_${k}_source.csv 2>/dev/null
rm -f gpurun_out/r2t2_$k.ncu-rep
tail -n 2 gpurun_out/r2t2_$k.log
