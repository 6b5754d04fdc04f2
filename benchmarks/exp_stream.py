"""Sweep over kernel sets (warp-specialised k_forward / k_linback, thread-per-problem k_forward_tp / k_linback_tp, build
variants), drain compaction on / off and slot counts, on BASELINE configs[1] (acrobot, T = 101): `batches` x 4096
problems streamed through S solver slots.
Usage: python benchmarks/exp_stream.py [cases=default,default:tp,...] [slots=14208,28416] [batches=20]
A case is  <build variant>[:tp|:tpback|:tpfwd|:ring][:nocompact]  -- which thread-per-problem kernels are forced on
(the others are forced off), and whether the drain compaction is disabled."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ilqr_b200  # noqa: F401
from bench import synth_inputs
from ilqr_b200 import build, capi, problems

kv = dict(a.split("=", 1) for a in sys.argv[1:])
cases = kv.get("cases", "default:ring:nocompact,default:ring,default:tp:nocompact,default:tp").split(",")
slots = [int(s) for s in kv.get("slots", "14208,18944,28416,37888").split(",")]
batches = int(kv.get("batches", "20"))
name = kv.get("model", "acrobot")
T = int(kv.get("T", "101" if name == "acrobot" else "51"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import inputs
model = getattr(problems, name)()
h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, 4096)
xs, us = [], []
for s in range(batches):
    if name == "acrobot":
        x1, ubar = synth_inputs(4096, T, seed=s)
    else:
        _, x1, ubar = inputs(name, 4096, T, seed=s)
    xs.append(h.rollout(x1, ubar)); us.append(ubar)
h.close()
xbar, ubar = np.concatenate(xs), np.concatenate(us)
n = xbar.shape[0]
dx, du = torch.from_numpy(xbar).cuda(), torch.from_numpy(ubar).cuda()
ref = None
for case in cases:
    parts = case.split(":")
    variant = "" if parts[0] == "default" else parts[0]
    never = str(1 << 40)
    os.environ["ILQR_TP_MIN_BLOCKS"] = "0" if ("tp" in parts or "tpback" in parts) else never
    os.environ["ILQR_FT_MIN_BLOCKS"] = "0" if ("tp" in parts or "tpfwd" in parts) else never
    if "nodense" in parts:
        os.environ["ILQR_LB_DENSE_MIN_BLOCKS"] = never
    else:
        os.environ.pop("ILQR_LB_DENSE_MIN_BLOCKS", None)
    os.environ["ILQR_FWD_TMA"] = "1" if "tma" in parts else ("2" if "sring" in parts else "0")
    if "nocompact" in parts:
        os.environ["ILQR_COMPACT_MIN_BLOCKS"] = never
    else:
        os.environ.pop("ILQR_COMPACT_MIN_BLOCKS", None)
    for sl in slots:
        hh = capi.Handle(build.model_library(model, variant=variant), T, model.n, model.m, model.p, model.cs, model.ct, sl, history_cap=1)
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
        hh.set_profiling(True)
        run()
        torch.cuda.synchronize()
        c = hh.get_counters()
        hh.set_profiling(False)
        kms, kl = [float(v) for v in c["kernel_ms"]], [int(v) for v in c["kernel_launches"]]
        print(json.dumps({"model": name, "case": case, "slots": sl, "problems": n, "ms": round(ms, 2), "solves_per_s": round(n / ms * 1e3),
                          "ticks": ticks, "us_per_tick": round(1e3 * ms / max(ticks, 1), 1),
                          "fwd_us": round(1e3 * kms[0] / max(kl[0], 1), 1), "lin_us": round(1e3 * kms[1] / max(kl[1], 1), 1),
                          "back_us": round(1e3 * kms[2] / max(kl[2], 1), 1), "problem_ticks": int(c["problem_ticks"]),
                          "ns_per_problem_tick": {"fwd": round(1e6 * kms[0] / max(int(c["problem_ticks"]), 1), 2),
                                                  "back": round(1e6 * (kms[1] + kms[2]) / max(int(c["problem_ticks"]), 1), 2)},
                          "compactions": int(c["compactions"]), "same_bits_as_first": sig == ref}), flush=True)
        hh.close()
