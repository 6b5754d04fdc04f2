#!/bin/bash
# round 2, GPU call B: parity with k_forward_tp + drain compaction, full-fill kernel throughput, streamed-job sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
timeout 300 python benchmarks/exp_fullfill.py cases=default,lb6,default:tp,default:tpfwd > gpurun_out/r2b_fullfill.jsonl 2> gpurun_out/r2b_fullfill.err
timeout 600 python benchmarks/exp_stream.py > gpurun_out/r2b_stream.jsonl 2> gpurun_out/r2b_stream.err
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 30 python benchmarks/sanitize_driver.py models=car,acrobot > gpurun_out/r2b_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2b_sanitizer_$tool.txt
done
tail -15 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_fullfill.jsonl; tail -c 600 gpurun_out/r2b_fullfill.err; cut -c1-400 gpurun_out/r2b_stream.jsonl; tail -c 600 gpurun_out/r2b_stream.err
for t in memcheck racecheck; do tail -6 gpurun_out/r2b_sanitizer_$t.txt; done
