#!/bin/bash
# round 2, GPU call O: cooperative Jacobian linearisation for table-mode wide models, time-varying dimensions, JAC_CONST
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2o_pytest.log
tail -n 8 gpurun_out/r2o_pytest.log
ILQR_VARIANT=rltimers C4_BATCH=296 timeout 200 python profiles/prof_c4.py 2>&1 | grep -E "phase cycles" | tail -1
timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2o_bench_c4.json 2> gpurun_out/r2o_bench.err
for f in r2o_bench_c4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$f", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "parity", (d.get("parity") or {}).get("ok"), "roofline", r["kernel"][:12], round(r["achieved"],2), round(r["frac"],3), {k:round(v["us_per_launch"]/1e3,3) for k,v in r["kernels"].items()})
except Exception as e: print("ERR", e)
PY
done
tail -n 3 gpurun_out/r2o_bench.err
