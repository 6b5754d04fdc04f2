"""Sweep: BASELINE configs[1] (acrobot, T = 101) streamed through S solver slots for several build variants
(line-search trials per launch).  Usage: python benchmarks/exp_slots.py [variants=,ls1] [slots=4096,8192,16384] [batches=10]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import ilqr_b200  # noqa: F401
from bench import synth_inputs
from ilqr_b200 import build, capi, problems

kv = dict(a.split("=", 1) for a in sys.argv[1:])
variants = kv.get("variants", ",ls1,ls4").split(",")
slots = [int(s) for s in kv.get("slots", "4096,8192,16384").split(",")]
batches = int(kv.get("batches", "10"))
T = 101
model = problems.acrobot()
h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, 4096)
xs, us = [], []
for s in range(batches):
    x1, ubar = synth_inputs(4096, T, seed=s)
    xs.append(h.rollout(x1, ubar)); us.append(ubar)
h.close()
xbar, ubar = np.concatenate(xs), np.concatenate(us)
n = xbar.shape[0]
dx, du = torch.from_numpy(xbar).cuda(), torch.from_numpy(ubar).cuda()
ref = None
for v in variants:
    for sl in slots:
        hh = capi.Handle(build.model_library(model, variant=v), T, model.n, model.m, model.p, model.cs, model.ct, sl, history_cap=1)
        st = torch.cuda.Stream(); hh.set_stream(st.cuda_stream)
        ox, ou = torch.empty_like(dx), torch.empty_like(du)
        it = torch.zeros(n, dtype=torch.int32, device="cuda")
        run = lambda: hh.solve_stream(n, dx.data_ptr(), du.data_ptr(), 0, ox.data_ptr(), ou.data_ptr(), it.data_ptr(), 0, 0, 0)
        run()
        c0 = hh.get_counters()["ticks"]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); run(); e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ticks = hh.get_counters()["ticks"] - c0
        sig = (ox.cpu().numpy().tobytes(), it.cpu().numpy().tobytes())
        if ref is None:
            ref = sig
        print(json.dumps({"variant": v or "default", "slots": sl, "problems": n, "ms": ms, "solves_per_s": n / ms * 1e3,
                          "ticks": ticks, "us_per_tick": 1e3 * ms / max(ticks, 1), "iterations_mean": float(it.float().mean()),
                          "same_bits_as_first": sig == ref}), flush=True)
        hh.close()
