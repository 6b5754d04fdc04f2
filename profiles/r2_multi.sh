#!/bin/bash
# round 2, 8-GPU call: BASELINE configs at their stated scale (one process per GPU, torchrun, NCCL gather through the C ABI)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
run() { # name, extra bench args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
      > gpurun_out/r2_bench_${name}_n$N.json 2> gpurun_out/r2_bench_${name}_n$N.err
  echo "== $name rc=$?"; tail -c 300 gpurun_out/r2_bench_${name}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${name}_n$N.json").read().strip().splitlines()[-1])
    print(json.dumps({k:d.get(k) for k in ("metric","value","n_gpus","ms_per_step","per_rank_ms","gather_ms","ticks_per_step","slot_fill","compactions","iterations_per_problem","converged_frac","parity")}))
    print("e2e", d.get("e2e")); print("roofline", {k:(d.get("roofline") or {}).get(k) for k in ("kernel","achieved","frac")})
except Exception as e:
    print("ERR", e)
PY
}
run c2 --steps 20 --warmup 3
run c3 --config c3 --steps 20 --warmup 3
run c5 --config c5 --steps 1000 --warmup 3 --no-cpu-baseline
