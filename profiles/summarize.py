"""Turn the raw captures of profiles/capture.sh (gpurun_out/<prefix>_*) into the committed summaries under profiles/.
    python profiles/summarize.py [prefix=r2] [slots=37888]"""
import sys
import collections
import csv
import json
import os
import re
import shutil
import subprocess

def kernel_base_name(name):
    """'void k_forward<3>(Params)' -> 'k_forward' (return type, namespace, template arguments and parameters dropped)"""
    m = re.search(r"([\w:]+)\s*(<[^()]*>)?\s*\(", name)
    return (m.group(1) if m else name).split("::")[-1]


kv = dict(a.split("=", 1) for a in sys.argv[1:])
PFX = kv.get("prefix", "r2")
SLOTS = int(kv.get("slots", "37888"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

for name in os.listdir(G):  # small text artefacts of this round's capture travel as they are
    if name.startswith(PFX + "_") and name.rsplit(".", 1)[-1] in ("json", "jsonl", "csv", "txt") and os.path.getsize(os.path.join(G, name)) < 2_000_000:
        shutil.copy(os.path.join(G, name), os.path.join(P, name))

# launch list -> per-kernel shares
rows = [r for r in csv.reader(open(os.path.join(G, PFX + "_launches.csv"))) if len(r) > 5]
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[idx["Metric Value"]])
    except ValueError:
        continue
    if r[idx["Metric Unit"]] == "ns":
        v /= 1000.0
    k = kernel_base_name(r[idx["Kernel Name"]])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
launch_summary = {k: {"launches": v[0], "total_us": v[1], "avg_us": v[1] / v[0], "share": v[1] / tot} for k, v in agg.items()}

# full capture -> selected metrics per kernel launch
raw_csv = os.path.join(G, PFX + "_prof_raw.csv")  # written by capture.sh on the GPU box (the .ncu-rep itself does not travel)
if os.path.exists(raw_csv):
    raw = open(raw_csv).read()
else:
    raw = subprocess.run(["ncu", "-i", os.path.join(G, PFX + "_prof.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["Kernel Name", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_ldgsts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct", "lts__t_bytes.sum", "smsp__cycles_active.avg", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
full = []
for r in rr[2:]:
    d = dict(zip(h, r))
    full.append({k: (d.get(k), dict(zip(h, units)).get(k)) for k in want if k in d})
traffic = {}
for e in full:
    k = kernel_base_name(e["Kernel Name"][0])
    def mb(x):
        v, u = x
        v = float(v)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic.setdefault(k, []).append(mb(e["dram__bytes_read.sum"]) + mb(e["dram__bytes_write.sum"]))
traffic = {k: sum(v) / len(v) for k, v in traffic.items()}
json.dump({"launch_list_shares": launch_summary, "full_capture": full, "dram_bytes_per_launch": traffic},
          open(os.path.join(P, PFX + "_ncu_summary.json"), "w"), indent=1)
json.dump({"note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, {SLOTS} slots, all problems iterating (profiles/capture.sh)",
           "config": "c2", "problems_per_launch": SLOTS, "dram_bytes_per_launch": traffic}, open(os.path.join(P, PFX + "_traffic.json"), "w"), indent=1)
print(json.dumps(launch_summary, indent=1))
print(traffic)

# config 4: DRAM bytes per launch of the three wide-model kernels (one working launch each, batch 1024)
c4 = {}
for kname in ("k_forward_wp", "k_linearize_jac", "k_backward"):
    f = os.path.join(G, f"{PFX}_prof_c4_{kname}_raw.csv")
    if not os.path.exists(f):
        continue
    rr4 = list(csv.reader(open(f)))
    h4, u4 = rr4[0], rr4[1]
    d4 = dict(zip(h4, rr4[2]))
    uu = dict(zip(h4, u4))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    c4[kname] = float(d4["dram__bytes_read.sum"]) * scale[uu["dram__bytes_read.sum"]] + float(d4["dram__bytes_write.sum"]) * scale[uu["dram__bytes_write.sum"]]
if c4:
    json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, batch 1024, a launch in which every problem iterates (profiles/capture.sh, profiles/prof_c4.py)",
               "config": "c4", "problems_per_launch": 1024, "dram_bytes_per_launch": c4}, open(os.path.join(P, PFX + "_traffic_c4.json"), "w"), indent=1)
    print(c4)
