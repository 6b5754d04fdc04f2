"""Shared fixtures: seeded synthetic inputs per model and comparison helpers."""
import numpy as np

import ilqr_b200  # noqa: F401
from ilqr_b200 import problems

HIST_KEYS = ("cost", "gradient_norm", "max_violation", "step_size", "outer", "status")


def inputs(name: str, B: int, T: int, seed: int = 0):
    rng = np.random.default_rng(seed)
    if name == "particle":
        model = problems.particle()
        x1 = np.zeros((B, 2))
        ubar = 0.1 * rng.standard_normal((B, T - 1, 1))
    elif name == "acrobot":
        model = problems.acrobot()
        x1 = 0.1 * rng.standard_normal((B, 4))
        ubar = rng.standard_normal((B, T - 1, 1))
    elif name == "car":
        model = problems.car()
        x1 = 0.05 * rng.uniform(-1, 1, (B, 3))
        ubar = np.tile(1e-2 * np.array([1.0, 0.1]), (B, T - 1, 1))
    elif name == "pendulum":
        model = problems.pendulum()
        x1 = np.tile(np.array([np.pi, 0.0]), (B, 1)) + 0.3 * rng.standard_normal((B, 2))
        ubar = 0.1 * rng.standard_normal((B, T - 1, 1))
    else:
        raise KeyError(name)
    return model, x1, ubar


def lq_inputs(B: int, T: int, n: int = 8, m: int = 2, seed: int = 0):
    """BASELINE config 4 in small: per-problem scaled dynamics and a sinusoid reference, both entering
    through the per-step parameter vector w_t = [s_b; r_{b,t}]."""
    rng = np.random.default_rng(seed)
    model = problems.lq_tracking(n, m)
    w = np.zeros((B, T, 2 * n))
    w[:, :, :n] = rng.uniform(-1, 1, (B, 1, n))
    tt = np.arange(T)[None, :, None]
    w[:, :, n:] = np.sin(0.3 * tt + rng.uniform(0, 6, (B, 1, n)))
    x1 = rng.standard_normal((B, n))
    ubar = 0.3 * rng.standard_normal((B, T - 1, m))
    return model, x1, ubar, w


def assert_same_solution(got, ref, bitwise=True):
    """got/ref: dicts with stats, history, x, u.  The north-star tolerance is: identical
    iteration counts, per-iteration cost / violation history to 1e-9 relative, trajectories
    to 1e-7; the engine and the C oracle share an arithmetic contract, so by default we
    demand bit equality, which implies the tolerance."""
    np.testing.assert_array_equal(got["stats"]["iterations"], ref["stats"]["iterations"])
    np.testing.assert_array_equal(got["stats"]["status"], ref["stats"]["status"])
    for k in ("cost", "max_violation", "gradient_norm", "step_size"):
        np.testing.assert_allclose(got["history"][k], ref["history"][k], rtol=1e-9, atol=0, err_msg=k)
    np.testing.assert_array_equal(got["history"]["outer"], ref["history"]["outer"])
    np.testing.assert_array_equal(got["history"]["status"], ref["history"]["status"])
    np.testing.assert_allclose(got["x"], ref["x"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(got["u"], ref["u"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(got["stats"]["objective"], ref["stats"]["objective"], rtol=1e-9)
    np.testing.assert_allclose(got["stats"]["max_violation"], ref["stats"]["max_violation"], rtol=1e-9, atol=1e-300)
    if bitwise:
        for k in ("cost", "max_violation", "gradient_norm", "step_size"):
            np.testing.assert_array_equal(got["history"][k], ref["history"][k], err_msg=k + " (bitwise)")
        np.testing.assert_array_equal(got["x"], ref["x"])
        np.testing.assert_array_equal(got["u"], ref["u"])
        np.testing.assert_array_equal(got["stats"]["flags"], ref["stats"]["flags"])


def collect(solver_like, cap=None):
    x, u = solver_like.get_trajectory()
    return dict(stats=solver_like.get_stats(), history=solver_like.get_history(cap), x=x, u=u)
