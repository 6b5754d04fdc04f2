"""Generate the golden fixtures under tests/golden/ (committed, small).

Julia is not available in this image, so the fixtures cannot come from IterativeLQR.jl
itself; they are produced by the two CPU oracles of this repo on seeded inputs:
  *_py.npz : oracle/ilqr_oracle.py (literal numpy + LAPACK restatement), tolerance target
  *_c.npz  : oracle/ilqr_oracle.c  (the arithmetic contract), bit-exact target for the engine
Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import inputs  # noqa: E402
from oracle.c_oracle import CModelFns, COracle  # noqa: E402
from oracle.ilqr_oracle import OracleSolver  # noqa: E402

CASES = {"particle": (11, 4, 11), "car": (51, 2, 12), "acrobot": (51, 2, 13), "pendulum": (31, 3, 14)}


def run_c(name):
    T, B, seed = CASES[name]
    model, x1, ubar = inputs(name, B, T, seed)
    co = COracle(model, T, B)
    xbar = co.rollout(x1, ubar)
    co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
    st, h = co.get_stats(), co.get_history()
    x, u = co.get_trajectory()
    n = int(st["iterations"].max())
    return dict(xbar=xbar, iterations=st["iterations"], objective=st["objective"], max_violation=st["max_violation"],
                cost=h["cost"][:, :n], viol=h["max_violation"][:, :n], gnorm=h["gradient_norm"][:, :n],
                step=h["step_size"][:, :n], outer=h["outer"][:, :n], x=x, u=u)


def run_py(name):
    T, B, seed = CASES[name]
    model, x1, ubar = inputs(name, B, T, seed)
    co = COracle(model, T, B)
    xbar = co.rollout(x1, ubar)
    fns = CModelFns(model)
    its, cost, viol, xs, us = [], [], [], [], []
    for b in range(B):
        dyn, obj, con = fns.as_reference_objects(T)
        s = OracleSolver(dyn, obj, con)
        s.initialize_controls([ubar[b, t] for t in range(T - 1)])
        s.initialize_states([xbar[b, t] for t in range(T)])
        s.solve()
        its.append(s.iterations[0])
        cost.append([r["cost"] for r in s.history]); viol.append([r["max_violation"] for r in s.history])
        xs.append(np.array(s.nominal_states)); us.append(np.array(s.nominal_actions[:-1]))
    n = max(its)
    pad = lambda rows: np.array([r + [0.0] * (n - len(r)) for r in rows])  # noqa: E731
    return dict(iterations=np.array(its, np.int32), cost=pad(cost), viol=pad(viol), x=np.array(xs), u=np.array(us))


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, f"{name}_c.npz"), **run_c(name))
        np.savez_compressed(os.path.join(HERE, f"{name}_py.npz"), **run_py(name))
        print(name, "done")
