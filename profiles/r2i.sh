#!/bin/bash
# round 2, GPU call I: time-varying stage functions, augmented_lagrangian_callback!, full suite, sanitizer on the new default kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2i_pytest.log
tail -n 25 gpurun_out/r2i_pytest.log
timeout 420 compute-sanitizer --tool racecheck --print-limit 30 python benchmarks/sanitize_driver.py models=car,acrobot > gpurun_out/r2i_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2i_sanitizer_racecheck.txt
timeout 420 compute-sanitizer --tool memcheck --print-limit 30 python benchmarks/sanitize_driver.py models=particle,car > gpurun_out/r2i_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2i_sanitizer_memcheck.txt
timeout 420 compute-sanitizer --tool synccheck --print-limit 30 python benchmarks/sanitize_driver.py models=acrobot > gpurun_out/r2i_sanitizer_synccheck.txt 2>&1; echo "synccheck rc=$?" >> gpurun_out/r2i_sanitizer_synccheck.txt
for t in racecheck memcheck synccheck; do tail -n 3 gpurun_out/r2i_sanitizer_$t.txt; done
