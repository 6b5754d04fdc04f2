#!/bin/bash
# round 2, 2-GPU check of the final code: the driver's own launch form for c2 (own arm and reference arm) and c4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
for cfg in c2 c4; do
  extra=""; [ $cfg = c4 ] && extra="--config c4 --steps 3"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra \
      > gpurun_out/r2_bench_${cfg}_n$N.json 2> gpurun_out/r2_bench_${cfg}_n$N.err
  echo "== $cfg rc=$?"; tail -c 300 gpurun_out/r2_bench_${cfg}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${cfg}_n$N.json").read().strip().splitlines()[-1])
    print(json.dumps({k:d.get(k) for k in ("metric","value","n_gpus","ms_per_step","per_rank_ms","gather_ms","scaling")}))
    print("e2e", (d.get("e2e") or {}).get("value"), "roofline", {k:(d.get("roofline") or {}).get(k) for k in ("kernel","achieved","frac","traffic")}, "parity", (d.get("parity") or {}).get("ok"))
except Exception as e:
    print("ERR", e)
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_n$N.json 2> gpurun_out/r2_bench_reference_n$N.err
echo "== reference rc=$?"; tail -c 400 gpurun_out/r2_bench_reference_n$N.json
