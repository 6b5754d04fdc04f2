#!/bin/bash
# round 2, GPU call X: Riccati kernel with helper warps (12 warps: 8 tile warps, factorisation warp, 3 solve warps), spine-first order
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "wide or large or constant_jac or lq" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2x_pytest.log
tail -n 5 gpurun_out/r2x_pytest.log
for v in "" rl_nohelpers; do
  ILQR_VARIANT=$v timeout 600 python bench.py --config c4 --steps 3 --no-cpu-baseline > gpurun_out/r2x_bench_c4_$v.json 2>> gpurun_out/r2x_bench.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2x_bench_c4_$v.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("variant '$v'", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "parity", (d.get("parity") or {}).get("ok"), round(r["frac"],3), {k:round(x["us_per_launch"]/1e3,3) for k,x in r["kernels"].items()})
except Exception as e: print("ERR", e)
PY
done
tail -n 3 gpurun_out/r2x_bench.err
