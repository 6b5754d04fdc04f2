#!/bin/bash
# round 2, GPU call Y: compute-sanitizer on the wide-model kernels; config-4 bench line with the final bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python benchmarks/sanitize_driver.py models=wide > gpurun_out/r2_sanitizer_wide_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver" gpurun_out/r2_sanitizer_wide_$tool.txt | tail -6
done
timeout 900 python -m pytest tests -m gpu -x -q -k "wide or large or constant_jac or lq" 2>&1 | tail -3; timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2y_bench.err
tail -c 600 gpurun_out/r2_bench_c4.json; tail -n 3 gpurun_out/r2y_bench.err
