#!/bin/bash
# round 2, GPU call P: per-kernel launch list of one config-4 batch solve (ncu durations are serialised / cold-cache: relative shares)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2p_c4_launches.csv python profiles/prof_c4.py > gpurun_out/r2p_c4.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2p_c4_launches.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",","")) * (1e-3 if r[ui]=="ns" else 1.0 if r[ui] in ("us","usecond") else 1e3 if r[ui]=="ms" else 1e-3))
    except Exception: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{k:60s} n={len(v):4d} mean={sum(v)/len(v):10.1f} us max={max(v):10.1f}")
PY
