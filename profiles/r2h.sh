#!/bin/bash
# round 2, GPU call H: cp.async-fed shared forward ring (sring): parity, throughput; config 4 parity on a fresh handle
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sring or kernel_variants" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2h_pytest.log
timeout 300 python benchmarks/exp_fullfill.py cases=default,default:sring batch=4096,14208,28416,37888 > gpurun_out/r2h_fullfill.jsonl 2> gpurun_out/r2h_fullfill.err
timeout 300 python benchmarks/exp_stream.py cases=default:ring,default:ring:sring slots=14208,37888 > gpurun_out/r2h_stream.jsonl 2> gpurun_out/r2h_stream.err
timeout 300 python benchmarks/exp_stream.py model=car cases=default:ring,default:ring:sring slots=2048,9472 batches=8 > gpurun_out/r2h_stream_car.jsonl 2>> gpurun_out/r2h_stream.err
timeout 600 python bench.py --config c4 --steps 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2h_bench.err
timeout 600 python bench.py --config c4 --steps 2 2>> gpurun_out/r2h_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4', d['value'], d['parity'])"
tail -n 4 gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_fullfill.jsonl; cut -c1-400 gpurun_out/r2h_stream.jsonl gpurun_out/r2h_stream_car.jsonl; tail -c 300 gpurun_out/r2h_stream.err gpurun_out/r2h_bench.err
