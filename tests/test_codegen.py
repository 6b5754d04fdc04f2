"""Generated model functions vs independent derivatives -- the reference's unit tests restated:
/root/reference/test/dynamics.jl:1-52 (pendulum Jacobians vs ForwardDiff, 1e-8),
test/objective.jl:1-40 (closed-form cost gradients), test/constraints.jl:1-43 (box constraint)."""
import numpy as np
import pytest

import ilqr_b200
from ilqr_b200 import Constraint, Cost, Dynamics, dot, problems, vcat
from oracle.c_oracle import CModelFns


def _fd_jac(f, x, eps=1e-6):
    x = np.asarray(x, float)
    y0 = np.asarray(f(x))
    J = np.zeros((y0.size, x.size))
    for j in range(x.size):
        d = np.zeros_like(x)
        d[j] = eps
        J[:, j] = (np.asarray(f(x + d)) - np.asarray(f(x - d))) / (2 * eps)
    return J


def test_dynamics_pendulum():  # test/dynamics.jl:21-51
    d = Dynamics(problems.pendulum_discrete, 2, 1)
    x1, u1 = np.ones(2), np.ones(1)

    def f(x, u):
        h = 0.1
        return np.array([x[0] + h * x[1], x[1] + h * (u[0] - 9.81 * np.sin(x[0]) - 0.1 * x[1])])

    d.evaluate(d.evaluate_cache, x1, u1, None)
    assert np.linalg.norm(d.evaluate_cache - f(x1, u1)) < 1e-8          # :32
    d.jacobian_state(d.jacobian_state_cache, x1, u1, None)
    assert np.linalg.norm(d.jacobian_state_cache - _fd_jac(lambda x: f(x, u1), x1)) < 1e-8   # :37
    d.jacobian_action(d.jacobian_action_cache, x1, u1, None)
    assert np.linalg.norm(d.jacobian_action_cache - _fd_jac(lambda u: f(x1, u), u1)) < 1e-8  # :43


def test_objective_closed_form():  # test/objective.jl:6-40
    ot = Cost(lambda x, u: dot(x, x) + 0.1 * dot(u, u), 2, 1)
    oT = Cost(lambda x, u: 10.0 * dot(x, x), 2, 0)
    x1, u1 = np.ones(2), np.ones(1)
    ot.evaluate(ot.evaluate_cache, x1, u1, None)
    assert abs(ot.evaluate_cache[0] - (2.0 + 0.1)) < 1e-8
    ot.gradient_state(ot.gradient_state_cache, x1, u1, None)
    assert np.linalg.norm(ot.gradient_state_cache - 2.0 * x1) < 1e-8    # :26
    ot.gradient_action(ot.gradient_action_cache, x1, u1, None)
    assert np.linalg.norm(ot.gradient_action_cache - 0.2 * u1) < 1e-8   # :27
    oT.gradient_state(oT.gradient_state_cache, x1, [], None)
    assert np.linalg.norm(oT.gradient_state_cache - 20.0 * x1) < 1e-8   # :28
    ot.hessian_state_state(ot.hessian_state_state_cache, x1, u1, None)
    assert np.allclose(ot.hessian_state_state_cache, 2 * np.eye(2))


def test_constraints_box():  # test/constraints.jl:13-43
    ct = Constraint(lambda x, u: vcat(-1.0 * np.ones(2) - x, x - np.ones(2)), 2, 1, indices_inequality=range(4))
    cT = Constraint(lambda x, u: x, 2, 0)
    x1, u1 = np.array([0.3, -2.0]), np.ones(1)
    ct.evaluate(ct.evaluate_cache, x1, u1, None)
    assert np.linalg.norm(ct.evaluate_cache - np.concatenate([-1 - x1, x1 - 1])) < 1e-8   # :27
    cT.evaluate(cT.evaluate_cache, x1, [], None)
    assert np.linalg.norm(cT.evaluate_cache - x1) < 1e-8                                  # :28
    ct.jacobian_state(ct.jacobian_state_cache, x1, u1, None)
    assert np.allclose(ct.jacobian_state_cache, np.vstack([-np.eye(2), np.eye(2)]))       # :40-43
    ct.jacobian_action(ct.jacobian_action_cache, x1, u1, None)
    assert np.allclose(ct.jacobian_action_cache, 0.0)
    assert ct.indices_inequality == [0, 1, 2, 3] and cT.indices_inequality == []


@pytest.mark.parametrize("name", ["particle", "pendulum", "car", "acrobot"])
def test_emitted_c_matches_symbolic_and_finite_differences(name):
    """The C that the kernels inline (compiled here for the host) against the lambdified
    sympy expressions and central differences."""
    model = getattr(problems, name)()
    fns = CModelFns(model)
    rng = np.random.default_rng(3)
    n, m = model.n, model.m
    for _ in range(5):
        x, u = rng.standard_normal(n), rng.standard_normal(m)
        y, fx, fu = fns.dyn(x, u)
        d = model.dynamics
        ye, fxe, fue = np.zeros(n), np.zeros((n, n)), np.zeros((n, m))
        d.evaluate(ye, x, u, None); d.jacobian_state(fxe, x, u, None); d.jacobian_action(fue, x, u, None)
        np.testing.assert_allclose(y, ye, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(fx, fxe, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(fu, fue, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(fx, _fd_jac(lambda xx: fns.dyn(xx, u)[0], x), atol=2e-7)   # 1e-8-class check of test/dynamics.jl
        np.testing.assert_allclose(fu, _fd_jac(lambda uu: fns.dyn(x, uu)[0], u), atol=2e-7)
        g, gx, gu, gxx, guu, gux = fns.cost(False, x, u)
        np.testing.assert_allclose(gx, _fd_jac(lambda xx: [fns.cost(False, xx, u)[0]], x)[0], atol=1e-5, rtol=1e-6)
        np.testing.assert_allclose(gu, _fd_jac(lambda uu: [fns.cost(False, x, uu)[0]], u)[0], atol=1e-5, rtol=1e-6)
        np.testing.assert_allclose(gxx, _fd_jac(lambda xx: fns.cost(False, xx, u)[1], x), atol=1e-5, rtol=1e-6)
        np.testing.assert_allclose(guu, _fd_jac(lambda uu: fns.cost(False, x, uu)[2], u), atol=1e-5, rtol=1e-6)
        np.testing.assert_allclose(gux, _fd_jac(lambda xx: fns.cost(False, xx, u)[2], x), atol=1e-5, rtol=1e-6)
        gT, gxT, _, gxxT, _, _ = fns.cost(True, x)
        np.testing.assert_allclose(gxT, _fd_jac(lambda xx: [fns.cost(True, xx)[0]], x)[0], atol=1e-4, rtol=1e-6)
        if model.cs:
            c, cx, cu = fns.con(False, x, u)
            np.testing.assert_allclose(cx, _fd_jac(lambda xx: fns.con(False, xx, u)[0], x), atol=1e-6)
            np.testing.assert_allclose(cu, _fd_jac(lambda uu: fns.con(False, x, uu)[0], u), atol=1e-6)
        if model.ct:
            c, cx, _ = fns.con(True, x)
            np.testing.assert_allclose(cx, _fd_jac(lambda xx: fns.con(True, xx)[0], x), atol=1e-6)


def test_header_is_deterministic():
    a = problems.acrobot().header
    from ilqr_b200 import api, codegen
    m2 = api.Model("acrobot", Dynamics(problems.acrobot_discrete, 4, 1), problems.acrobot().cost_stage,
                   problems.acrobot().cost_terminal, problems.acrobot().con_stage, problems.acrobot().con_terminal)
    assert m2.header == a and codegen.header_hash(a) == problems.acrobot().hash


def test_inequality_indices_validated():
    with pytest.raises(ValueError):
        Constraint(lambda x, u: x, 2, 0, indices_inequality=[2])


def test_grouped_trig_evaluation_is_bit_identical_to_single_calls(monkeypatch):
    """codegen._emit_trig_group: independent sin/cos arguments share one range test and run the branch-free body
    back to back.  Every argument must still see exactly the operations of ilqr_sincos -- also when one argument
    of a group is huge (whole group takes the fallback) and when a sin/cos feeds another one (two levels)."""
    import sympy as sp
    from ilqr_b200 import api, codegen
    sin, cos = sp.sin, sp.cos

    def f(x, u):
        a = sin(x[0]) * cos(x[1]) + sin(x[0] + x[1]) + cos(x[2] * u[0])
        b = sin(cos(x[0]) + x[2]) + cos(x[1]) * u[0]            # second level: argument depends on a cosine
        return [x[0] + 0.1 * a, x[1] + 0.1 * b, x[2] + 0.05 * sin(x[2]) / (2.0 + cos(x[0]))]

    def build(group):
        monkeypatch.setattr(codegen, "TRIG_GROUP", group)
        d = Dynamics(f, 3, 1)
        c = Cost(lambda x, u: dot(x, x) + dot(u, u), 3, 1)
        cT = Cost(lambda x, u: dot(x, x), 3, 0)
        model = api.Model(f"trig{group}", d, c, cT, Constraint(), Constraint())
        assert model.header  # emitted now, while TRIG_GROUP is patched
        return model

    grouped, single = build(6), build(1)
    assert "ilqr_sincos_small" in grouped.header and "ilqr_sincos_small" not in single.header
    fg, fs = CModelFns(grouped), CModelFns(single)
    rng = np.random.default_rng(11)
    for k in range(200):
        x, u = rng.standard_normal(3) * 3.0, rng.standard_normal(1)
        if k % 4 == 1:
            x[rng.integers(3)] *= 1e6       # beyond the Cody-Waite range: the group falls back to ilqr_sincos
        if k % 4 == 2:
            x *= 1e13
        for got, ref in zip(fg.dyn(x, u), fs.dyn(x, u)):
            np.testing.assert_array_equal(got, ref)
