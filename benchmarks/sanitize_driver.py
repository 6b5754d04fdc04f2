"""Small batches through the three modes of the engine (ilqr_solve, ilqr_solve_stream, ilqr_mpc_run), with both
decompositions of the backward tick, for compute-sanitizer:
    compute-sanitizer --tool racecheck|memcheck|synccheck python benchmarks/sanitize_driver.py [models=particle,car,acrobot]
Sizes are tiny on purpose (the tools slow kernels down by 10-100x); the mbarrier rings of k_linback, the hand ring
without an "empty" handshake and k_refill's concurrent side branch are all exercised."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import ilqr_b200  # noqa: F401
from common import inputs
from ilqr_b200 import build, capi

kv = dict(a.split("=", 1) for a in sys.argv[1:])
names = kv.get("models", "particle,car,acrobot").split(",")
if "wide" in names:
    # the wide-model kernels (n = 64, m = 16 dense plant): bulk copies + mbarriers in k_backward / k_forward_wp, named barriers
    # between the factorisation, solve and tile warps in every schedule variant, the cooperative Jacobian kernel
    from common import lq_inputs
    names.remove("wide")
    Bw, Tw = int(kv.get("Bw", "3")), int(kv.get("Tw", "6"))
    model, x1, ubar, w = lq_inputs(Bw, Tw, 64, 16, seed=11)
    # (the experimental 12-warp variant "rl_helpers" passes memcheck and racecheck; synccheck objects to its CTA-wide barrier being
    # reached from two code sites -- tile warps and helper warps -- so it is not in the default list)
    for variant in kv.get("wide_variants", ",rl_fwarp,nojacconst,nodmma").split(","):
        o = capi.default_options()
        o.max_iterations = 2
        o.objective_tolerance = 0.0
        o.lagrangian_gradient_tolerance = 0.0
        h = capi.Handle(build.model_library(model, variant=variant), Tw, model.n, model.m, model.p, model.cs, model.ct, Bw, options=o, history_cap=8)
        h.set_parameters(w)
        xbar = h.rollout(x1, ubar)
        h.initialize_controls(ubar); h.initialize_states(xbar); h.solve()
        st = h.get_stats()
        h.mpc_step()
        print(f"sanitize_driver: lq64x16 variant '{variant}': iterations {int(st['iterations'].min())}..{int(st['iterations'].max())}, "
              f"flags {int(st['flags'].max())}, mpc step ok", flush=True)
        h.close()
T, B, NS = int(kv.get("T", "9")), int(kv.get("B", "64")), int(kv.get("n", "96"))
for name in names:
    for tp, tma in (("0", "0"), (str(1 << 40), "0"), (str(1 << 40), "1"), (str(1 << 40), "2")):
        os.environ["ILQR_TP_MIN_BLOCKS"] = tp   # k_linback_tp on / off
        os.environ["ILQR_FWD_TMA"] = tma        # k_forward: private rings / one ring per CTA fed by bulk copies (1) or cp.async (2)
        os.environ["ILQR_COMPACT_MIN_BLOCKS"] = "1"  # the drain compaction runs even on these two-block grids
        model, x1, ubar = inputs(name, NS, T, seed=3)
        o = capi.default_options()
        o.max_iterations = 4
        o.max_dual_updates = 2
        h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B, options=o, history_cap=8)
        xbar = h.rollout(x1[:B], ubar[:B])
        h.initialize_controls(ubar[:B]); h.initialize_states(xbar); h.solve()
        it_batch = h.get_stats()["iterations"].copy()
        # stream NS problems through B slots
        xs = np.concatenate([xbar, h.rollout(np.resize(x1[B:], (B, model.n)), np.resize(ubar[B:], (B, T - 1, model.m)))[:NS - B]])
        dx, du = torch.from_numpy(xs).cuda(), torch.from_numpy(ubar).cuda()
        ox, ou = torch.zeros_like(dx), torch.zeros_like(du)
        it = torch.zeros(NS, dtype=torch.int32, device="cuda")
        h.solve_stream(NS, dx.data_ptr(), du.data_ptr(), 0, ox.data_ptr(), ou.data_ptr(), it.data_ptr(), 0, 0, 0)
        torch.cuda.synchronize()
        assert np.array_equal(it.cpu().numpy()[:B], it_batch), "streamed iteration counts differ from the batch solve"
        # receding horizon
        h.initialize_controls(ubar[:B]); h.initialize_states(xbar); h.solve()
        au = torch.zeros((3, B, model.m), dtype=torch.float64, device="cuda")
        xn = torch.zeros((3, B, model.n), dtype=torch.float64, device="cuda")
        h.mpc_run(3, au.data_ptr(), xn.data_ptr(), 0)
        torch.cuda.synchronize()
        c = h.get_counters()
        print(f"sanitize_driver: {name} tp_min_blocks={tp} fwd_tma={tma}: batch iterations {int(it_batch.min())}..{int(it_batch.max())}, "
              f"stream ok, mpc ok, {c['ticks']} ticks, {c['compactions']} compaction(s)", flush=True)
        h.close()
