#!/bin/bash
# round 2, GPU call L: DMMA Riccati (wide-model path): parity, config 4 bench with and without DMMA, ncu of the Riccati kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wide or large" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2l_pytest.log
tail -n 15 gpurun_out/r2l_pytest.log
timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2l_bench_c4.json 2> gpurun_out/r2l_bench.err
ILQR_VARIANT=nodmma timeout 900 python bench.py --config c4 --steps 3 --no-cpu-baseline > gpurun_out/r2l_bench_c4_nodmma.json 2>> gpurun_out/r2l_bench.err
for f in r2l_bench_c4 r2l_bench_c4_nodmma; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$f", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("bitwise_equal"), "roofline", round(r["achieved"],2), round(r["frac"],3), {k:round(v["us_per_launch"]/1e3,2) for k,v in r["kernels"].items()})
except Exception as e: print("ERR", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_backward -s 1 -c 1 -o gpurun_out/r2_prof_c4 python profiles/prof_c4.py > gpurun_out/r2_prof_c4.log 2>&1
ncu -i gpurun_out/r2_prof_c4.ncu-rep --page raw --csv > gpurun_out/r2_prof_c4_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2_prof_c4_raw.csv"))); h=rows[0]
want=["Kernel Name","gpu__time_duration.sum","launch__registers_per_thread","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__warps_active.avg.pct_of_peak_sustained_active","l1tex__lsu_writeback_active_mem_lg.avg.pct_of_peak_sustained_elapsed","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d=dict(zip(h,r)); print({k:d.get(k) for k in want if k in d})
PY
tail -n 3 gpurun_out/r2l_bench.err gpurun_out/r2_prof_c4.log
