#!/bin/bash
cd "$(dirname "$0")/.."
bash profiles/capture.sh
for f in r2_bench_n1 r2_bench_c3 r2_bench_c5_100steps r2_bench_reference; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    keep={k:d.get(k) for k in ("metric","value","ms_per_step","ticks_per_step","slot_fill","compactions","gather_ms","iterations_per_problem","converged_frac","gpu_launches","parity","cpu_baseline","lockstep_batches")}
    keep["e2e"]=d.get("e2e"); r=d.get("roofline") or {}
    keep["roofline"]={k:r.get(k) for k in ("kernel","achieved","peak","frac","traffic")}
    keep["kernels"]={k:{kk:round(vv,3) if isinstance(vv,float) else vv for kk,vv in v.items()} for k,v in (r.get("kernels") or {}).items()}
    print(json.dumps(keep))
except Exception as e:
    print("ERR", e)
PY
done
