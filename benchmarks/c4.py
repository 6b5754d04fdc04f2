#!/usr/bin/env python
"""BASELINE config 4: synthetic dense linear-quadratic tracking, T=256, n=64, m=16, batch 1024 (wide-model path:
CTA-per-problem shared-memory Riccati kernel).  Parity of the first problems against the C oracle, then timing.
The plug-in for this model takes ~35 minutes to compile (5000-statement generated functions); it is cached
under iterativelqr.jl_b200/_build/ and travels with the snapshot."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import ilqr_b200  # noqa: F401
from common import lq_inputs
from ilqr_b200 import build, capi
from oracle.c_oracle import COracle

T, B, NCHECK = 256, int(os.environ.get("C4_BATCH", "1024")), int(os.environ.get("C4_CHECK", "16"))
t0 = time.time()
model, x1, ubar, w = lq_inputs(B, T, 64, 16, seed=0)
if os.environ.get("C4_MODEL") == "banded":  # fast-compiling stand-in with the same shapes (problems.lq_banded)
    from ilqr_b200 import problems
    model = problems.lq_banded(64, 16)
ubar[:] = 0.0  # u = 0, x = rollout
print("model traced in %.1f s" % (time.time() - t0), flush=True)
h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B, history_cap=16)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
h.set_parameters(w)
xbar = h.rollout(x1, ubar)

def solve():
    h.initialize_controls(ubar); h.initialize_states(xbar); h.solve()

solve()
torch.cuda.synchronize()
h.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st); h.initialize_controls(ubar); h.initialize_states(xbar); h.solve(); e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
c = h.get_counters()
stats = h.get_stats()
xg, ug = h.get_trajectory()
hist = h.get_history(16)
out = {"config": f"C4 dense LQ tracking T={T} n=64 m=16 p=128 batch {B}", "ms_per_batch_solve": ms, "solves_per_s": B / ms * 1e3,
       "ticks": c["ticks"], "iterations_mean": float(stats["iterations"].mean()), "iterations_max": int(stats["iterations"].max()),
       "kernel_ms": {k: float(v) for k, v in zip(("forward", "linearize", "backward"), c["kernel_ms"])},
       "kernel_launches": [int(v) for v in c["kernel_launches"]], "flags": np.unique(stats["flags"]).tolist()}
# Riccati FP64 accounting: 4n^3+10n^2m+6nm^2+m^3/3 flops per step
n, m = 64, 16
flops_step = 4 * n**3 + 10 * n * n * m + 6 * n * m * m + m**3 / 3
launches = int(c["kernel_launches"][2])
pt = int(c["problem_ticks"])  # problems actually worked on, summed over the launches
out["riccati"] = {"flops_per_problem_pass": flops_step * (T - 1), "problem_passes": pt, "ms_total": float(c["kernel_ms"][2]),
                  "launches": launches, "tflops": flops_step * (T - 1) * pt / (float(c["kernel_ms"][2]) * 1e-3) / 1e12,
                  "algorithmic_gbs": (8 * (2 * n * n + 2 * n * m + m * m + n + m) + 8 * (m * n + 2 * m + n)) * (T - 1) * pt
                                     / (float(c["kernel_ms"][2]) * 1e-3) / 1e9}
print(json.dumps(out), flush=True)

# parity of the first problems against the oracle (fresh solver each)
t0 = time.time()
co = COracle(model, T, NCHECK, history_cap=16)
co.set_parameters(w[:NCHECK])
xo = co.rollout(x1[:NCHECK], ubar[:NCHECK])
print("rollout bit-equal:", bool(np.array_equal(xo, xbar[:NCHECK])))
h2 = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, NCHECK, history_cap=16)
h2.set_parameters(w[:NCHECK]); h2.initialize_controls(ubar[:NCHECK]); h2.initialize_states(xbar[:NCHECK]); h2.solve()
co.initialize_controls(ubar[:NCHECK]); co.initialize_states(xo); co.solve()
so, ho = co.get_stats(), co.get_history(16)
sg, hg = h2.get_stats(), h2.get_history(16)
xoo, uoo = co.get_trajectory(); xgg, ugg = h2.get_trajectory()
res = {"oracle_seconds": time.time() - t0, "oracle_threads": co.max_threads(), "checked_problems": NCHECK,
       "iterations_equal": bool(np.array_equal(sg["iterations"], so["iterations"])),
       "history_cost_bit_equal": bool(np.array_equal(hg["cost"], ho["cost"])),
       "history_cost_max_rel": float(np.max(np.abs(hg["cost"] - ho["cost"]) / np.maximum(np.abs(ho["cost"]), 1e-300))),
       "x_bit_equal": bool(np.array_equal(xgg, xoo)), "u_bit_equal": bool(np.array_equal(ugg, uoo)),
       "x_max_abs_diff": float(np.abs(xgg - xoo).max()), "u_max_abs_diff": float(np.abs(ugg - uoo).max()),
       "iterations": sg["iterations"].tolist(), "cpu_solves_per_s": NCHECK / max(time.time() - t0, 1e-9)}
print(json.dumps(res), flush=True)
