#!/bin/bash
# round 2, GPU call Z: synccheck on the wide-model kernels (default list of variants), full GPU test suite on the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python benchmarks/sanitize_driver.py models=wide > gpurun_out/r2_sanitizer_wide_synccheck.txt 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|sanitize_driver" gpurun_out/r2_sanitizer_wide_synccheck.txt | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2z_pytest.log
tail -n 6 gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
