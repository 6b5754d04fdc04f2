#!/bin/bash
# round 2, GPU call AA: k_forward_wp -- conflict-light row staging for the sweep, pipelined nominal copy, unrolled dynamics rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "wide or large or constant_jac or lq" > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2aa_pytest.log
tail -n 4 gpurun_out/r2aa_pytest.log
timeout 600 python bench.py --config c4 --steps 3 > gpurun_out/r2aa_bench_c4.json 2>> gpurun_out/r2aa_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2aa_bench_c4.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("c4", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "parity", (d.get("parity") or {}).get("ok"), round(r["frac"],3), {k:round(x["ms_per_launch_all_problems_working"],2) for k,x in r["kernels"].items()})
except Exception as e: print("ERR", e)
PY
tail -n 3 gpurun_out/r2aa_bench.err
