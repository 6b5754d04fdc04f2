#!/bin/bash
# Round-1 evidence capture (run under gpurun on ONE B200):   bash profiles/capture.sh
# Produces gpurun_out/r1_* files; profiles/summarize.py turns them into the committed summaries.

cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r1_clocks.csv &
SMI=$!
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
kill $SMI
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench_n1.err
# every launch of the same command with its device time (cold-cache, serialised: compare SHARES).  Launches 200-600 of
# the process fall into the warm-up streaming job (3 x 4096 problems through 12288 slots) while every slot is busy.
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r1_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_under_ncu.log 2>&1
# full capture of the two hot kernels at the bench's operating point: 14208 slots, all problems iterating
PROF_BATCH=14208 PROF_MAX_ITERS=60 ncu --set full --clock-control none --import-source on -k regex:k_ -s 100 -c 4 -o gpurun_out/r1_prof \
    python profiles/prof_driver.py > gpurun_out/r1_prof.log 2>&1
python benchmarks/configs.py c3 c5 > gpurun_out/r1_configs.jsonl 2>> gpurun_out/r1_bench_n1.err
python benchmarks/exp_slots.py variants=,ls1,ls4 slots=4096,8192,12288,14208,16384 > gpurun_out/r1_slots_sweep.jsonl 2>> gpurun_out/r1_bench_n1.err
./profiles/microbench/fp64_latency > gpurun_out/r1_fp64_latency.txt 2>&1
tail -c 400 gpurun_out/r1_bench_n1.json
