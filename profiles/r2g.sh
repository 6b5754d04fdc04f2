#!/bin/bash
# round 2, GPU call G: wide-model path with the per-problem Hessian accumulator + dense-table dynamics; config 4 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2g_pytest.log
timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench.err
tail -n 4 gpurun_out/r2g_pytest.log
for f in r2g_bench_c4; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    keep={k:d.get(k) for k in ("metric","value","ms_per_step","ticks_per_step","slot_fill","iterations_per_problem","converged_frac","parity","cpu_baseline")}
    keep["e2e"]=d.get("e2e"); r=d.get("roofline") or {}
    keep["roofline"]={k:r.get(k) for k in ("kernel","achieved","peak","unit","frac","traffic")}
    keep["kernels"]={k:{kk:round(vv,3) if isinstance(vv,float) else vv for kk,vv in v.items()} for k,v in (r.get("kernels") or {}).items()}
    print(json.dumps(keep))
except Exception as e:
    print("ERR", e)
PY
done
tail -n 5 gpurun_out/r2g_bench.err
