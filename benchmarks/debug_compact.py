"""Debugging aid for the drain compaction: one streamed job against the oracle, with the engine logging every move
(ILQR_COMPACT_DEBUG=1); prints which problems differ.  python benchmarks/debug_compact.py [name T slots n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

os.environ["ILQR_COMPACT_DEBUG"] = "1"
os.environ.setdefault("ILQR_COMPACT_MIN_BLOCKS", "1")
os.environ.setdefault("ILQR_TP_MIN_BLOCKS", str(1 << 40))
import ilqr_b200  # noqa: F401
from common import inputs
from ilqr_b200 import build, capi
from oracle.c_oracle import COracle

name, T, slots, n = (sys.argv[1:] + ["acrobot", "31", "512", "1500"][len(sys.argv) - 1:])[:4]
T, slots, n = int(T), int(slots), int(n)
model, x1, ubar = inputs(name, n, T, seed=31)
co = COracle(model, T, n, history_cap=1)
xbar = co.rollout(x1, ubar)
co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
so = co.get_stats()
h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, slots, history_cap=1)
dx, du = torch.from_numpy(xbar).cuda(), torch.from_numpy(ubar).cuda()
ox, ou = torch.zeros_like(dx), torch.zeros_like(du)
it = torch.zeros(n, dtype=torch.int32, device="cuda")
h.solve_stream(n, dx.data_ptr(), du.data_ptr(), 0, ox.data_ptr(), ou.data_ptr(), it.data_ptr(), 0, 0, 0)
torch.cuda.synchronize()
got = it.cpu().numpy()
bad = np.nonzero(got != so["iterations"])[0]
print("mismatching problems:", bad.tolist(), "engine", got[bad].tolist(), "oracle", so["iterations"][bad].tolist(), "counters", h.get_counters())
