"""Kernel throughput with every slot busy: one lock-step batch of B acrobot problems (T = 101), 30 iterations,
CUDA events around every launch.  Separates what a kernel can sustain from the fill of a streamed job.
Usage: python benchmarks/exp_fullfill.py [cases=default,default:tp,...] [batch=4736,9472,...]
A case is <build variant>[:tp|:tpback|:tpfwd] -- which of the thread-per-problem kernels are forced on (others off)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ilqr_b200  # noqa: F401
from bench import algorithmic_bytes, synth_inputs
from ilqr_b200 import build, capi, problems

kv = dict(a.split("=", 1) for a in sys.argv[1:])
cases = kv.get("cases", "default,lb6,default:tp").split(",")
batches = [int(s) for s in kv.get("batch", "4736,9472,14208,18944,28416,37888").split(",")]
T = 101
model = problems.acrobot()
ab = algorithmic_bytes(model, T)
NEVER = str(1 << 40)
for case in cases:
    variant, _, mode = case.partition(":")
    variant = "" if variant == "default" else variant
    os.environ["ILQR_TP_MIN_BLOCKS"] = "0" if mode in ("tp", "tpback") else NEVER
    os.environ["ILQR_FT_MIN_BLOCKS"] = "0" if mode in ("tp", "tpfwd") else NEVER
    os.environ["ILQR_FWD_TMA"] = {"tma": "1", "sring": "2"}.get(mode, "0")
    for B in batches:
        x1, ubar = synth_inputs(B, T, seed=B)
        o = capi.default_options()
        o.max_iterations = 30
        o.max_dual_updates = 1
        h = capi.Handle(build.model_library(model, variant=variant), T, model.n, model.m, model.p, model.cs, model.ct, B, options=o, history_cap=1)
        xbar = h.rollout(x1, ubar)
        for rep in range(2):
            h.initialize_controls(ubar); h.initialize_states(xbar)
            if rep == 1:
                h.set_profiling(True)
            h.solve()
        c = h.get_counters()
        kms, kl = [float(v) for v in c["kernel_ms"]], [int(v) for v in c["kernel_launches"]]
        pt = int(c["problem_ticks"])
        fill = pt / (c["ticks"] * B)
        f_us, b_us = 1e3 * kms[0] / kl[0], 1e3 * (kms[1] + kms[2]) / kl[2]
        print(json.dumps({"case": case, "batch": B, "ticks": int(c["ticks"]), "fill": round(fill, 3),
                          "fwd_us": round(f_us, 1), "back_us": round(b_us, 1),
                          "fwd_ns_per_problem_tick": round(1e6 * kms[0] / pt, 2), "back_ns_per_problem_tick": round(1e6 * (kms[1] + kms[2]) / pt, 2),
                          "back_algorithmic_gbs": round(ab["backward"] * pt / ((kms[1] + kms[2]) * 1e-3) / 1e9),
                          "fwd_algorithmic_gbs": round(ab["forward"] * pt / (kms[0] * 1e-3) / 1e9)}), flush=True)
        h.close()
