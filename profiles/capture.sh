#!/bin/bash
# Round-2 evidence capture (run under gpurun on ONE B200):   bash profiles/capture.sh
# Produces gpurun_out/r2_* files; profiles/summarize.py turns them into the committed summaries.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DWITH_M16N8K8 -o profiles/microbench/fp64_peak profiles/microbench/fp64_peak.cu 2>/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench/fp64_latency profiles/microbench/fp64_latency.cu 2>/dev/null
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2_clocks.csv &
SMI=$!
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
kill $SMI
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --config c3 > gpurun_out/r2_bench_c3.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --config c5 --steps 100 > gpurun_out/r2_bench_c5_100steps.json 2>> gpurun_out/r2_bench_n1.err
./profiles/microbench/fp64_peak > gpurun_out/r2_fp64_peak.txt 2>&1
./profiles/microbench/fp64_latency > gpurun_out/r2_fp64_latency.txt 2>&1
# every launch of the same command with its device time (cold-cache, serialised: compare SHARES).  Launches 300-700 of
# the process fall into the warm-up streaming job while every slot is busy.
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
# full capture of the two hot kernels at the bench's operating point: 37888 slots, all problems iterating
PROF_BATCH=37888 PROF_MAX_ITERS=60 ncu --set full --clock-control none --import-source on -k regex:k_ -s 100 -c 4 -o gpurun_out/r2_prof \
    python profiles/prof_driver.py > gpurun_out/r2_prof.log 2>&1
ncu -i gpurun_out/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2>/dev/null
# config 4 (wide-model path): bench line, launch list and full captures of its three kernels (one working launch each)
python bench.py --config c4 --steps 3 > gpurun_out/r2_bench_c4.json 2>> gpurun_out/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_c4_launches.csv python profiles/prof_c4.py > gpurun_out/r2_c4_launches.log 2>&1
for k in k_forward_wp k_linearize_jac k_backward; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r2_prof_c4_$k python profiles/prof_c4.py > gpurun_out/r2_prof_c4_$k.log 2>&1
  ncu -i gpurun_out/r2_prof_c4_$k.ncu-rep --page raw --csv > gpurun_out/r2_prof_c4_${k}_raw.csv 2>/dev/null
  rm -f gpurun_out/r2_prof_c4_$k.ncu-rep
done
rm -f gpurun_out/r2_prof.ncu-rep
tail -c 1500 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_prof.log | tail -3
