#!/bin/bash
# round 2, GPU call F: full-row stores (no partial-sector fills), dense C4 model, bench parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2f_pytest.log
timeout 300 python benchmarks/exp_fullfill.py cases=default,default:tpback batch=14208,28416,37888 > gpurun_out/r2f_fullfill.jsonl 2> gpurun_out/r2f_fullfill.err
timeout 300 python benchmarks/exp_stream.py cases=default:ring slots=14208,28416,37888 > gpurun_out/r2f_stream.jsonl 2> gpurun_out/r2f_stream.err
timeout 300 python bench.py > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --config c4 --steps 3 > gpurun_out/r2f_bench_c4.json 2>> gpurun_out/r2f_bench.err
PROF_BATCH=37888 PROF_MAX_ITERS=60 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 100 -c 2 -o gpurun_out/r2f_prof python profiles/prof_driver.py > gpurun_out/r2f_prof.log 2>&1
ncu -i gpurun_out/r2f_prof.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    d=dict(zip(h,r)); print(d['Kernel Name'], 'us', d['gpu__time_duration.sum'], 'rd', d['dram__bytes_read.sum'], 'wr', d['dram__bytes_write.sum'], 'tex_rd_sectors', d.get('lts__t_sectors_srcunit_tex_op_read.sum'), 'tex_wr_sectors', d.get('lts__t_sectors_srcunit_tex_op_write.sum'))
"
tail -n 4 gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_fullfill.jsonl; cut -c1-420 gpurun_out/r2f_stream.jsonl
for f in r2f_bench_c2 r2f_bench_c4; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    keep={k:d.get(k) for k in ("metric","value","ms_per_step","ticks_per_step","slot_fill","compactions","iterations_per_problem","converged_frac","parity","cpu_baseline")}
    keep["e2e"]=d.get("e2e"); r=d.get("roofline") or {}
    keep["roofline"]={k:r.get(k) for k in ("kernel","achieved","peak","unit","frac","traffic")}
    keep["kernels"]={k:{kk:round(vv,3) if isinstance(vv,float) else vv for kk,vv in v.items()} for k,v in (r.get("kernels") or {}).items()}
    print(json.dumps(keep))
except Exception as e:
    print("ERR", e)
PY
done
tail -n 5 gpurun_out/r2f_bench.err
