#!/bin/bash
# round 2, 8-GPU call on the final code: BASELINE config 3 (car, 16384 problems per step) with the 37888-slot default, and config 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
for cfg in c3 c2; do
  extra=""; [ $cfg = c3 ] && extra="--config c3"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra --no-cpu-baseline \
      > gpurun_out/r2_bench_${cfg}_n${N}_final.json 2> gpurun_out/r2_bench_${cfg}_n${N}_final.err
  echo "== $cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${cfg}_n${N}_final.json").read().strip().splitlines()[-1])
    print(json.dumps({k:d.get(k) for k in ("metric","value","n_gpus","ms_per_step","per_rank_ms","gather_ms","slot_fill","compactions")}))
    print("e2e", (d.get("e2e") or {}).get("value"), "parity", (d.get("parity") or {}).get("ok"))
except Exception as e:
    print("ERR", e)
PY
done
