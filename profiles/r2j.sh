#!/bin/bash
# round 2, GPU call J: speculative trials without trajectory writes; full suite; bench c2/c3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2j_pytest.log
tail -n 6 gpurun_out/r2j_pytest.log
timeout 300 python benchmarks/exp_fullfill.py cases=default batch=4096,14208,28416,37888 > gpurun_out/r2j_fullfill.jsonl 2> gpurun_out/r2j_fullfill.err
timeout 300 python benchmarks/exp_stream.py cases=default:ring:sring slots=14208,28416,37888 > gpurun_out/r2j_stream.jsonl 2> gpurun_out/r2j_stream.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_c2.json 2> gpurun_out/r2j_bench.err
timeout 300 python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2j_bench_c3.json 2>> gpurun_out/r2j_bench.err
cat gpurun_out/r2j_fullfill.jsonl; cut -c1-420 gpurun_out/r2j_stream.jsonl
for f in r2j_bench_c2 r2j_bench_c3; do python - <<PY
import json
d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("$f", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ticks/step", d["ticks_per_step"], {k:(round(v["ns_per_problem_tick"],2), round(v["frac_of_hbm_peak"],3)) for k,v in r["kernels"].items()})
PY
done
tail -n 3 gpurun_out/r2j_bench.err
