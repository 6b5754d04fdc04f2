#!/bin/bash
# round 2, GPU call R: factorisation and solves under the Qxx contraction, out-of-line unrolled triangular solves, flat single-term Jacobian tables, L1 bypass for scattered loads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2r_pytest.log
tail -n 8 gpurun_out/r2r_pytest.log
ILQR_VARIANT=rltimers C4_BATCH=296 timeout 200 python profiles/prof_c4.py 2>&1 | grep -E "phase cycles" | tail -1
ILQR_VARIANT=rltimers_nooverlap C4_BATCH=296 timeout 200 python profiles/prof_c4.py 2>&1 | grep -E "phase cycles" | tail -1
timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2r_bench_c4.json 2> gpurun_out/r2r_bench.err
for f in r2r_bench_c4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$f", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "parity", (d.get("parity") or {}).get("ok"), "roofline", r["kernel"][:12], round(r["achieved"],2), round(r["frac"],3), {k:round(v["us_per_launch"]/1e3,3) for k,v in r["kernels"].items()})
except Exception as e: print("ERR", e)
PY
done
tail -n 3 gpurun_out/r2r_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2r_c4_launches.csv python profiles/prof_c4.py > gpurun_out/r2r_c4.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2r_c4_launches.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",","")) * (1e-3 if r[ui]=="ns" else 1.0 if r[ui] in ("us","usecond") else 1e3 if r[ui]=="ms" else 1e-3))
    except Exception: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:6]: print(f"{k:60s} n={len(v):4d} mean={sum(v)/len(v):10.1f} us max={max(v):10.1f}")
PY
