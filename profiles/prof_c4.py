"""One config-4 batch solve for ncu:  ncu --set full -k regex:k_backward -s 1 -c 1 -o gpurun_out/r1_prof_c4 python profiles/prof_c4.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ilqr_b200  # noqa
from common import lq_inputs
from ilqr_b200 import build, capi
T, B = 256, int(os.environ.get("C4_BATCH", "1024"))
model, x1, ubar, w = lq_inputs(B, T, 64, 16, seed=0); ubar[:] = 0
if os.environ.get("C4_MODEL") == "banded":
    from ilqr_b200 import problems
    model = problems.lq_banded(64, 16)
h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B, history_cap=4)
h.set_parameters(w); xbar = h.rollout(x1, ubar); h.initialize_controls(ubar); h.initialize_states(xbar); h.solve()
print(h.get_counters()["ticks"], h.get_stats()["iterations"].mean())
