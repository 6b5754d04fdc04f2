"""One batched acrobot solve (BASELINE configs[1]) for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:k_ -s 300 -c 9 -o gpurun_out/prof python profiles/prof_driver.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ilqr_b200  # noqa: F401
from bench import synth_inputs
from ilqr_b200 import build, capi, problems

B = int(os.environ.get("PROF_BATCH", "4096"))
T = 101
model = problems.acrobot()
x1, ubar = synth_inputs(B, T)
h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B, history_cap=8)
xbar = h.rollout(x1, ubar)
h.initialize_controls(ubar)
h.initialize_states(xbar)
if os.environ.get("PROF_MAX_ITERS"):
    o = capi.default_options()
    o.max_iterations = int(os.environ["PROF_MAX_ITERS"])
    o.max_dual_updates = 2
    h.set_options(o)
h.solve()
print(h.get_counters(), int(h.get_stats()["iterations"].max()))
