#!/bin/bash
# round 2, GPU call C: parity after the ls_base fix + dense k_linback; streamed-job sweeps (acrobot, car)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
timeout 500 python benchmarks/exp_stream.py cases=default:ring:nocompact,default:ring,default:ring:nodense,default:tpback slots=9472,14208,18944,28416,37888 > gpurun_out/r2c_stream.jsonl 2> gpurun_out/r2c_stream.err
timeout 300 python benchmarks/exp_stream.py model=car cases=default:ring:nodense,default:ring,default:tpback slots=2048,4736,9472,18944 batches=8 > gpurun_out/r2c_stream_car.jsonl 2>> gpurun_out/r2c_stream.err
tail -8 gpurun_out/r2c_pytest.log; cut -c1-420 gpurun_out/r2c_stream.jsonl; cut -c1-420 gpurun_out/r2c_stream_car.jsonl; tail -c 600 gpurun_out/r2c_stream.err
