"""Flip-rate study (test infrastructure; SURVEY.md section 7: "report the flip rate").

How robust is "same iteration count, history to 1e-9, trajectories to 1e-7" to the things a different host
implementation of IterativeLQR.jl's algorithm legitimately differs in?  Three implementations of the same solve:

  A  oracle/ilqr_oracle.py, literal statement order, numpy matmul + LAPACK dpotrf/dpotrs, model functions =
     sympy-lambdified callables on libm sin/cos (the closest stand-in available here for the Julia package's own
     Symbolics-generated functions);
  B  the same literal oracle driven by the EMITTED C model functions (ilqr_sincos etc., what the kernels inline);
  C  oracle/ilqr_oracle.c, the arithmetic contract the CUDA engine reproduces bit for bit.

A vs C isolates "everything": trig implementation, expression evaluation order, BLAS/LAPACK vs fma chains;
B vs C isolates the linear-algebra order alone.  Output: one JSON document (profiles/r2_flip_rate.json).

    python oracle/flip_rate_study.py [n=256] [procs=6] [models=acrobot,car] > profiles/r2_flip_rate.json
"""
import json
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

HORIZON = {"acrobot": 101, "car": 51, "particle": 11, "pendulum": 31}


def _py_solve(args):
    name, T, b, n, lambdified = args
    from common import inputs
    from test_oracle import py_solver
    model, x1, ubar = inputs(name, n, T, seed=0)
    s = py_solver(model, T, x1[b], ubar[b], lambdified=lambdified)
    s.solve()
    x, u = s.get_trajectory()
    return dict(b=b, iterations=int(s.iterations[0]), cost=[r["cost"] for r in s.history],
                viol=[r["max_violation"] for r in s.history], alpha=[r["step_size"] for r in s.history],
                x=np.array(x), u=np.array(u))


def compare(py, c_stats, c_hist, xc, uc):
    n = len(py)
    mism, rel_cost, rel_viol, dx, du, alpha_diff = [], 0.0, 0.0, 0.0, 0.0, 0
    for r in py:
        b, it = r["b"], r["iterations"]
        itc = int(c_stats["iterations"][b])
        if it != itc:
            k = min(it, itc)
            first = next((i for i in range(k) if r["alpha"][i] != c_hist["step_size"][b, i]), k)
            mism.append(dict(problem=b, iterations_py=it, iterations_c=itc, first_step_size_difference_at=first))
            continue
        cc, vc = c_hist["cost"][b, :it], c_hist["max_violation"][b, :it]
        rel_cost = max(rel_cost, float(np.max(np.abs(np.array(r["cost"]) - cc) / np.abs(cc))))
        vp = np.array(r["viol"])
        nz = np.abs(vc) > 0
        if nz.any():
            rel_viol = max(rel_viol, float(np.max(np.abs(vp[nz] - vc[nz]) / np.abs(vc[nz]))))
        alpha_diff += int(np.sum(np.array(r["alpha"]) != c_hist["step_size"][b, :it]))
        dx = max(dx, float(np.max(np.abs(r["x"] - xc[b]))))
        du = max(du, float(np.max(np.abs(r["u"] - uc[b]))))
    return dict(problems=n, iteration_count_mismatches=len(mism), flip_rate=len(mism) / n, mismatched=mism[:16],
                among_matching=dict(cost_history_max_rel=rel_cost, violation_history_max_rel=rel_viol,
                                    step_size_records_differing=alpha_diff, x_max_abs=dx, u_max_abs=du),
                north_star=dict(history_rel=1e-9, trajectory_abs=1e-7,
                                met=bool(not mism and rel_cost <= 1e-9 and rel_viol <= 1e-9 and dx <= 1e-7 and du <= 1e-7)))


def main():
    kv = dict(a.split("=", 1) for a in sys.argv[1:])
    n, procs = int(kv.get("n", "256")), int(kv.get("procs", "6"))
    out = {"what": __doc__.split("\n\n")[1], "inputs": "tests/common.py inputs(name, n, T, seed=0): SURVEY 8d randomisation", "models": {}}
    from common import inputs
    from oracle.c_oracle import COracle
    for name in kv.get("models", "acrobot,car").split(","):
        T = HORIZON[name]
        model, x1, ubar = inputs(name, n, T, seed=0)
        co = COracle(model, T, n, history_cap=1000)
        xbar = co.rollout(x1, ubar)
        co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
        st, hist = co.get_stats(), co.get_history()
        xc, uc = co.get_trajectory()
        res = {"T": T, "iterations_mean": float(st["iterations"].mean()), "iterations_max": int(st["iterations"].max())}
        with mp.Pool(procs) as pool:
            for key, lamb in (("A_lambdified_libm_vs_C", True), ("B_emitted_functions_vs_C", False)):
                py = pool.map(_py_solve, [(name, T, b, n, lamb) for b in range(n)], chunksize=1)
                res[key] = compare(py, st, hist, xc, uc)
                print(name, key, json.dumps(res[key]["among_matching"]), "mismatches", res[key]["iteration_count_mismatches"], file=sys.stderr, flush=True)
        out["models"][name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
