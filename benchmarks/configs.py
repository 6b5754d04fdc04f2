#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (bench.py covers configs[1]).

    python benchmarks/configs.py [c3] [c5] [slots] [--mpc-steps S]

  c3     car with obstacle / box inequality constraints (AL), T=51, one GPU's shard of configs[2] (2048 problems)
  c5     receding-horizon MPC, acrobot T=101, one GPU's shard of configs[4] (1024 problems), S re-solves each
  slots  configs[1] streamed through 4096 / 8192 / 16384 slots (how throughput scales with problems in flight)
One JSON line per measurement.  Device-resident inputs, CUDA-event timing on the solve stream, 1 warm-up run.
"""
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
from bench import synth_inputs
from common import inputs
from ilqr_b200 import build, capi, problems


def timed(stream, fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); fn(); e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def stream_job(model, T, slots, xbar, ubar, label, extra=None):
    n = xbar.shape[0]
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, slots, history_cap=1)
    st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
    dx, du = torch.from_numpy(xbar).cuda(), torch.from_numpy(ubar).cuda()
    ox, ou = torch.empty_like(dx), torch.empty_like(du)
    it = torch.zeros(n, dtype=torch.int32, device="cuda"); mv = torch.zeros(n, dtype=torch.float64, device="cuda")
    run = lambda: h.solve_stream(n, dx.data_ptr(), du.data_ptr(), 0, ox.data_ptr(), ou.data_ptr(), it.data_ptr(), 0, 0, mv.data_ptr())
    run()
    c0 = h.get_counters()["ticks"]
    ms = timed(st, run)
    ticks = h.get_counters()["ticks"] - c0
    out = {"config": label, "problems": n, "slots": slots, "ms": ms, "solves_per_s": n / ms * 1e3, "ticks": ticks,
           "us_per_tick": 1e3 * ms / max(ticks, 1), "iterations_mean": float(it.float().mean()), "iterations_max": int(it.max()),
           "feasible_frac": float((mv <= 5e-3).double().mean())}
    out.update(extra or {})
    print(json.dumps(out), flush=True)
    h.close()


def c3():
    T, B = 51, 2048
    model, x1, ubar = inputs("car", 4 * B, T, seed=0)
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, 4 * B)
    xbar = h.rollout(x1, ubar); h.close()
    stream_job(model, T, B, xbar, ubar, "C3 car T=51 n=3 m=2 c_s=5 c_T=4: 4 x 2048 problems through 2048 slots (1 GPU's shard of 16384/8)")


def c5(steps):
    T, B = 101, 1024
    model = problems.acrobot()
    x1, ubar = synth_inputs(B, T, seed=2)
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B, history_cap=1)
    st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
    xbar = h.rollout(x1, ubar)
    h.initialize_controls(ubar); h.initialize_states(xbar)
    ms0 = timed(st, h.solve)
    tot = torch.zeros(B, dtype=torch.int32, device="cuda")
    c0 = h.get_counters()["ticks"]
    ms = timed(st, lambda: h.mpc_run(steps, 0, 0, tot.data_ptr()))
    ticks = h.get_counters()["ticks"] - c0
    print(json.dumps({"config": f"C5 acrobot MPC T=101: {B} problems (1 GPU's shard of 8192/8), {steps} warm-started re-solves each, asynchronous per problem",
                      "first_solve_ms": ms0, "ms": ms, "resolves_per_s": B * steps / ms * 1e3, "ms_per_mpc_step_per_problem_stream": ms / steps,
                      "ticks": ticks, "iterations_per_resolve_mean": float(tot.float().mean()) / steps,
                      "projected_1000_steps_s": ms / steps}), flush=True)
    h.close()


def slots():
    T = 101
    model = problems.acrobot()
    n = 4 * 4096
    xs, us = [], []
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, 4096)
    for s in range(4):
        x1, ubar = synth_inputs(4096, T, seed=s)
        xs.append(h.rollout(x1, ubar)); us.append(ubar)
    h.close()
    xbar, ubar = np.concatenate(xs), np.concatenate(us)
    for sl in (4096, 8192, 16384):
        stream_job(model, T, sl, xbar, ubar, f"C2 acrobot T=101: {n} problems through {sl} slots")


if __name__ == "__main__":
    args = sys.argv[1:]
    steps = 20
    if "--mpc-steps" in args:
        steps = int(args[args.index("--mpc-steps") + 1])
    if not args or "c3" in args:
        c3()
    if not args or "c5" in args:
        c5(steps)
    if not args or "slots" in args:
        slots()
