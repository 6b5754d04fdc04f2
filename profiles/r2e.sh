#!/bin/bash
# round 2, GPU call E: store-fill microbenchmark under ncu; k_forward_tma parity, sanitizer, throughput
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_write.sum --csv --log-file gpurun_out/r2_store_fill.csv ./profiles/microbench/store_fill > gpurun_out/r2_store_fill.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2e_pytest.log
timeout 300 python benchmarks/exp_fullfill.py cases=default,default:tma batch=4096,14208,28416,37888 > gpurun_out/r2e_fullfill.jsonl 2> gpurun_out/r2e_fullfill.err
timeout 300 python benchmarks/exp_stream.py cases=default:ring,default:ring:tma slots=14208,37888 > gpurun_out/r2e_stream.jsonl 2> gpurun_out/r2e_stream.err
timeout 300 python benchmarks/exp_stream.py model=car cases=default:ring,default:ring:tma slots=2048,9472 batches=8 > gpurun_out/r2e_stream_car.jsonl 2>> gpurun_out/r2e_stream.err
timeout 420 compute-sanitizer --tool racecheck --print-limit 30 python benchmarks/sanitize_driver.py models=car,acrobot > gpurun_out/r2e_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2e_sanitizer_racecheck.txt
timeout 420 compute-sanitizer --tool memcheck --print-limit 30 python benchmarks/sanitize_driver.py models=car > gpurun_out/r2e_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2e_sanitizer_memcheck.txt
grep -E "k_st|k_bulk" gpurun_out/r2_store_fill.csv | cut -d, -f5,13- ; tail -6 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_fullfill.jsonl; tail -c 400 gpurun_out/r2e_fullfill.err
cut -c1-400 gpurun_out/r2e_stream.jsonl gpurun_out/r2e_stream_car.jsonl; tail -c 400 gpurun_out/r2e_stream.err; tail -4 gpurun_out/r2e_sanitizer_racecheck.txt gpurun_out/r2e_sanitizer_memcheck.txt
