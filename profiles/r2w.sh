#!/bin/bash
# round 2, GPU call W: full GPU test suite on the final code, then the evidence capture (profiles/capture.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2w_pytest.log
tail -n 6 gpurun_out/r2w_pytest.log
timeout 2400 bash profiles/capture.sh 2>&1 | tail -12
tail -c 1200 gpurun_out/r2_bench_c4.json
