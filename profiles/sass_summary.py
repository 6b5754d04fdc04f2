"""cuobjdump -sass of a model plug-in -> per-kernel counts of the mnemonics that matter (TMA bulk copies, cp.async, mbarrier
ops, FP64 arithmetic, tensor-core ops).  python profiles/sass_summary.py [model=acrobot] > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilqr_b200  # noqa: F401
from ilqr_b200 import build, problems

kv = dict(a.split("=", 1) for a in sys.argv[1:])
name = kv.get("model", "acrobot")
model = {"lq64": lambda: problems.lq_tracking(64, 16)}.get(name, lambda: getattr(problems, name)())()
lib = build.model_library(model)
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip() or s
KEYS = ["UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "DFMA", "DMUL", "DADD", "DMMA", "MUFU.RCP64H", "LDG", "STG", "LDS", "STS", "BAR", "SHFL"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1)).split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
print(f"# {os.path.relpath(lib)}  (sm_100a SASS, static instruction counts per kernel)")
print(f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{k[:8]:>8s}" for k in KEYS))
for fn, c in counts.items():
    if "k_" not in fn:
        continue
    print(f"{fn[:58]:58s} {total[fn]:7d} " + " ".join(f"{c.get(k, 0):8d}" for k in KEYS))
