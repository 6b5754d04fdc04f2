#!/bin/bash
# round 2, GPU calls AB: A/B of the k_forward_wp experiments (row-wise staging, unroll factor of ilqr_dyn_part, pipelined nominal copy);
# the build variants named here (wp_*) were removed again after the measurement (profiles/README.md section 8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in wp_unroll4 wp_unroll16 wp_unroll40 wp_unroll80; do
  ILQR_VARIANT=$v timeout 600 python bench.py --config c4 --steps 3 --no-cpu-baseline > gpurun_out/r2ab_bench_c4_$v.json 2>> gpurun_out/r2ab_bench.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2ab_bench_c4_$v.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("variant '$v'", round(d["value"]), "ms/step", round(d["ms_per_step"],1), round(r["frac"],3), {k:round(x["ms_per_launch_all_problems_working"],2) for k,x in r["kernels"].items()})
except Exception as e: print("ERR", e)
PY
done
tail -n 3 gpurun_out/r2ab_bench.err
