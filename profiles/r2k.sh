#!/bin/bash
# round 2, GPU call K: warp-per-problem wide forward kernel: parity on the dense plant, config 4 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wide or large" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2k_pytest.log
tail -n 12 gpurun_out/r2k_pytest.log
timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench.err
ILQR_FWD_WP=0 timeout 900 python bench.py --config c4 --steps 3 --no-cpu-baseline > gpurun_out/r2k_bench_c4_nowp.json 2>> gpurun_out/r2k_bench.err
for f in r2k_bench_c4 r2k_bench_c4_nowp; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$f", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("bitwise_equal"), "roofline", round(r["achieved"],2), round(r["frac"],3), {k:round(v["us_per_launch"]/1e3,2) for k,v in r["kernels"].items()})
except Exception as e: print("ERR", e)
PY
done
tail -n 5 gpurun_out/r2k_bench.err
