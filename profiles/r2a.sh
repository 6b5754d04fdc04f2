#!/bin/bash
# round 2, GPU call A: parity of the new kernels, FP64 peaks, the linback sweep, compute-sanitizer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
timeout 120 ./profiles/microbench/fp64_peak > gpurun_out/r2a_fp64_peak.txt 2>&1
timeout 700 python benchmarks/exp_linback.py cases=nohacc,default,lb6,lb6k,default:tp,tp12:tp,nohacc:tp,ls1:tp \
    > gpurun_out/r2a_linback.jsonl 2> gpurun_out/r2a_linback.err
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench.err
ILQR_TP_MIN_BLOCKS=0 timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --slots 37888 > gpurun_out/r2a_bench_tp37888.json 2>> gpurun_out/r2a_bench.err
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 30 python benchmarks/sanitize_driver.py > gpurun_out/r2a_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2a_sanitizer_$tool.txt
done
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_fp64_peak.txt; cat gpurun_out/r2a_linback.jsonl | cut -c1-330; tail -c 300 gpurun_out/r2a_linback.err
for t in memcheck synccheck racecheck; do tail -4 gpurun_out/r2a_sanitizer_$t.txt; done
