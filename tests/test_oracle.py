"""The CPU oracles (test infrastructure) pinned against what the reference's own tests assert and
against each other.

Reference assertions restated: /root/reference/test/acrobot.jl:114 (terminal state within 5e-3),
test/car.jl:74-79 (all stage constraints, terminal equalities and the terminal inequality within
constraint_tolerance).  The reference's tests pin nothing else about the solve path."""
import os

import numpy as np
import pytest

from common import inputs
from oracle.c_oracle import CModelFns, COptions, COracle
from oracle.ilqr_oracle import Options, OracleSolver, rollout

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def py_solver(model, T, x1, ubar, lambdified=False, **attrs):
    if lambdified:
        dyn = [model.dynamics] * (T - 1)
        obj = [model.cost_stage] * (T - 1) + [model.cost_terminal]
        con = [model.con_stage] * (T - 1) + [model.con_terminal] if model.constrained else None
    else:
        dyn, obj, con = CModelFns(model).as_reference_objects(T)
    xbar = rollout(dyn, x1, list(ubar))
    s = OracleSolver(dyn, obj, con, options=Options(verbose=False))
    for k, v in attrs.items():
        setattr(s, k, v)
    s.initialize_controls(list(ubar))
    s.initialize_states(xbar)
    return s


def test_reference_acrobot_assertion_py_oracle():
    """test/acrobot.jl:88-114 with a seeded init (the reference's is an unseeded randn)."""
    model, x1, ubar = inputs("acrobot", 1, 51, seed=1)
    s = py_solver(model, 51, np.zeros(4), ubar[0], lambdified=True)
    s.solve()
    xT = np.array([np.pi, 0, 0, 0])
    assert np.max(np.abs(s.nominal_states[-1] - xT)) < 5e-3       # test/acrobot.jl:114
    assert s.iterations[0] == len(s.history) and not s.chol_fail


def test_reference_car_assertions_py_oracle():
    """test/car.jl:22-79, deterministic init."""
    model, _, ubar = inputs("car", 1, 51)
    s = py_solver(model, 51, np.zeros(3), ubar[0], lambdified=True)
    s.solve()
    x, u = s.get_trajectory()
    tol = s.options.constraint_tolerance
    c = np.zeros(5)
    for t in range(50):                                            # test/car.jl:74
        model.con_stage.evaluate(c, x[t], u[t], None)
        assert np.all(c <= tol)
    cT = np.zeros(4)
    model.con_terminal.evaluate(cT, x[50], [], None)
    assert cT[3] <= tol and np.all(np.abs(cT[:3]) <= tol)          # test/car.jl:77-79


def test_particle_quickstart_py_oracle():
    """README.md:39-88 / examples/particle.jl: reaches the goal."""
    model, x1, ubar = inputs("particle", 1, 11, seed=5)
    s = py_solver(model, 11, x1[0], ubar[0], lambdified=True)
    s.solve()
    assert np.max(np.abs(s.nominal_states[-1] - np.array([1.0, 0.0]))) < 5e-3


def test_quirks_change_the_history():
    """Q1 (Hessian accumulation) and Q2 (stale constraint values) are load-bearing: switching
    either off changes the iteration history, so an engine that 'fixes' them fails parity."""
    model, x1, ubar = inputs("particle", 1, 11, seed=5)
    runs = {}
    for q1 in (True, False):
        for q2 in (True, False):
            s = py_solver(model, 11, x1[0], ubar[0], accumulate_hessians=q1, stale_constraints=q2)
            s.solve()
            runs[(q1, q2)] = [r["cost"] for r in s.history]
    assert runs[(True, True)] != runs[(False, False)]
    assert runs[(True, True)] != runs[(True, False)]
    assert runs[(True, True)] != runs[(False, True)]


def test_lambdified_and_emitted_models_agree():
    """The oracle driven by sympy-lambdified callables vs by the emitted C (what the kernels inline)."""
    model, x1, ubar = inputs("car", 1, 51)
    a = py_solver(model, 51, x1[0], ubar[0], lambdified=True); a.solve()
    b = py_solver(model, 51, x1[0], ubar[0], lambdified=False); b.solve()
    assert a.iterations == b.iterations
    np.testing.assert_allclose([r["cost"] for r in a.history], [r["cost"] for r in b.history], rtol=1e-9)
    np.testing.assert_allclose([r["max_violation"] for r in a.history], [r["max_violation"] for r in b.history], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(np.array(a.nominal_states), np.array(b.nominal_states), rtol=0, atol=1e-7)


@pytest.mark.parametrize("name,T", [("particle", 11), ("car", 51), ("acrobot", 51), ("pendulum", 31)])
def test_c_oracle_matches_py_oracle(name, T):
    """oracle/ilqr_oracle.c (arithmetic contract) vs oracle/ilqr_oracle.py (literal numpy/LAPACK): same control flow
    (iteration counts, step sizes), histories and trajectories at the NORTH-STAR tolerance (cost / violation history to
    1e-9 relative, x and u to 1e-7).  profiles/r2_flip_rate.json holds the same comparison over 256 acrobot (T = 101) and
    256 car problems, with sympy-lambdified libm model functions as well: 0 iteration-count mismatches."""
    B = 4
    model, x1, ubar = inputs(name, B, T, seed=21)
    co = COracle(model, T, B)
    xbar = co.rollout(x1, ubar)
    co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
    st, h = co.get_stats(), co.get_history()
    xc, uc = co.get_trajectory()
    for b in range(B):
        s = py_solver(model, T, x1[b], ubar[b])
        np.testing.assert_array_equal(np.array(s.nominal_states), xbar[b])  # same emitted dynamics -> same rollout bits
        s.solve()
        n = s.iterations[0]
        assert n == st["iterations"][b]
        np.testing.assert_allclose([r["cost"] for r in s.history], h["cost"][b, :n], rtol=1e-9)
        np.testing.assert_allclose([r["max_violation"] for r in s.history], h["max_violation"][b, :n], rtol=1e-9, atol=1e-7)
        np.testing.assert_array_equal([r["step_size"] for r in s.history], h["step_size"][b, :n])
        np.testing.assert_array_equal([r["outer"] for r in s.history], h["outer"][b, :n])
        np.testing.assert_allclose(np.array(s.nominal_states), xc[b], rtol=0, atol=1e-7)
        np.testing.assert_allclose(np.array(s.nominal_actions[:-1]), uc[b], rtol=0, atol=1e-7)


@pytest.mark.parametrize("name", ["particle", "car", "acrobot", "pendulum"])
def test_c_oracle_reproduces_golden_bitwise(name):
    """Regression pin: committed fixtures (tests/golden/make_golden.py)."""
    from make_golden import run_c
    g = np.load(os.path.join(GOLD, f"{name}_c.npz"))
    r = run_c(name)
    for k in g.files:
        np.testing.assert_array_equal(r[k], g[k], err_msg=k)
    gp = np.load(os.path.join(GOLD, f"{name}_py.npz"))
    np.testing.assert_array_equal(gp["iterations"], g["iterations"])
    n = gp["cost"].shape[1]
    np.testing.assert_allclose(gp["cost"], g["cost"][:, :n], rtol=1e-9)
    np.testing.assert_allclose(gp["x"], g["x"], rtol=0, atol=1e-7)


def test_c_oracle_options_and_warm_start():
    """Option plumbing of the C oracle against the python oracle: line_search=:none (Q12),
    small iteration limits, reset_cache, and a warm-started second solve (Q2 state persists)."""
    model, x1, ubar = inputs("car", 1, 31, seed=4)
    for kw in (dict(line_search="none", max_iterations=5, max_dual_updates=3),
               dict(max_iterations=3, max_dual_updates=2),
               dict(reset_cache=True, max_dual_updates=3),
               dict(min_step_size=0.3),
               dict(max_iterations=0, max_dual_updates=2)):
        co = COracle(model, 31, 1, options=COptions.default(**{k: (int(v) if isinstance(v, bool) else v) for k, v in kw.items()}))
        xbar = co.rollout(x1, ubar)
        co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
        s = py_solver(model, 31, x1[0], ubar[0])
        for k, v in kw.items():
            setattr(s.options, k, v)
        s.solve()
        st, h = co.get_stats(), co.get_history()
        assert st["iterations"][0] == s.iterations[0], kw
        n = s.iterations[0]
        # records are indexed by data.iterations; reset_cache restarts that counter every inner solve,
        # so only the records of the last inner solve survive (the python oracle's list keeps all)
        hist = s.history[len(s.history) - n:]
        np.testing.assert_allclose([r["cost"] for r in hist], h["cost"][0, :n], rtol=1e-6, err_msg=str(kw))
        np.testing.assert_array_equal([r["step_size"] for r in hist], h["step_size"][0, :n])
        # warm start: solve!(solver, x, u) again on the same solver objects
        s.history.clear()
        s.solve(states=[xbar[0, t] for t in range(31)], actions=[ubar[0, t] for t in range(30)])
        co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
        st, h = co.get_stats(), co.get_history()
        assert st["iterations"][0] == s.iterations[0], ("warm", kw)
        n = s.iterations[0]
        hist = s.history[len(s.history) - n:]
        np.testing.assert_allclose([r["cost"] for r in hist], h["cost"][0, :n], rtol=1e-6, err_msg="warm " + str(kw))
