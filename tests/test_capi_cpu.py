"""CPU-side checks of the C-ABI boundary: libilqr_cuda.so loads and exports every symbol that
include/ilqr_cuda.h declares; the model plug-ins load; and WITHOUT a GPU the library fails
loudly instead of computing anything (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import ilqr_b200
from ilqr_b200 import build, capi, problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ilqr_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ilqr_[a-z_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 25
    L = ctypes.CDLL(build.front_library())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ilqr_cuda.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == declared, "python binding and header disagree"


def test_options_default_matches_reference_defaults():  # src/options.jl:1-14
    o = capi.default_options()
    assert (o.line_search, o.max_iterations, o.max_dual_updates, o.reset_cache) == (0, 100, 10, 0)
    assert (o.min_step_size, o.objective_tolerance, o.lagrangian_gradient_tolerance) == (1e-5, 1e-3, 1e-3)
    assert (o.constraint_tolerance, o.initial_constraint_penalty, o.scaling_penalty, o.max_penalty) == (5e-3, 1.0, 10.0, 1e8)
    assert o.constraint_norm == float("inf")
    from ilqr_b200 import Options
    assert bytes(Options().to_c())[:16] == bytes(o)[:16]


@pytest.mark.parametrize("name", ["particle", "car", "acrobot", "pendulum"])
def test_model_plugin_loads_and_reports_dims(name):
    model = getattr(problems, name)()
    path = build.model_library(model)
    assert capi.model_dims(path) == (model.n, model.m, model.p, model.cs, model.ct)


def test_create_rejects_bad_arguments():
    model = problems.particle()
    path = build.model_library(model)
    with pytest.raises(capi.IlqrError, match="cannot load model library"):
        capi.Handle("/nonexistent/libmodel.so", 11, 2, 1, 0, 0, 2, 4)
    with pytest.raises(capi.IlqrError, match="dimension mismatch"):
        capi.Handle(path, 11, 3, 1, 0, 0, 2, 4)
    with pytest.raises(capi.IlqrError, match="T must be"):
        capi.Handle(path, 1, 2, 1, 0, 0, 2, 4)
    with pytest.raises(capi.IlqrError, match="batch must be"):
        capi.Handle(path, 11, 2, 1, 0, 0, 2, 0)
    with pytest.raises(capi.IlqrError, match="not an ilqr model plug-in"):
        capi.Handle(build.front_library(), 11, 2, 1, 0, 0, 2, 4)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    model = problems.particle()
    with pytest.raises(capi.IlqrError, match="no CUDA device|CUDA"):
        capi.Handle(build.model_library(model), 11, 2, 1, 0, 0, 2, 4)
    from ilqr_b200 import Solver
    with pytest.raises(capi.IlqrError):
        Solver(model, T=11)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "iterativelqr.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "ilqr_oracle" not in text or f == "codegen.py" or "oracle/ilqr_oracle.c" in text, f


def test_time_varying_stage_functions_merge_into_one_model():
    """Distinct per-step Dynamics / Cost objects (src/solver.jl:28-30) are merged into ONE compiled stage function that
    selects its variant by a trailing parameter; the merged functions must reproduce every variant (checked through the
    emitted C, compiled for the host by the oracle's build recipe)."""
    from ilqr_b200 import Cost, Dynamics, dot
    from ilqr_b200.solver import _model_from_lists
    from oracle.c_oracle import CModelFns
    d1 = Dynamics(lambda x, u: [x[0] + 0.1 * x[1], x[1] + 0.1 * u[0]], 2, 1)
    d2 = Dynamics(lambda x, u: [x[0] + 0.1 * x[1], 0.9 * x[1] + 0.1 * u[0] - 0.05 * x[0] ** 3], 2, 1)
    c1 = Cost(lambda x, u: dot(x, x) + 0.1 * dot(u, u), 2, 1)
    c2 = Cost(lambda x, u: 2 * dot(x, x) + 0.1 * dot(u, u) + 0.01 * u[0] ** 4, 2, 1)
    cT = Cost(lambda x, u: 10 * dot(x, x), 2, 0)
    dyn, obj = [d1, d2, d1, d1, d2, d2], [c1, c1, c2, c1, c2, c1, cT]
    model, T, kinds, _ = _model_from_lists(dyn, obj, None)
    assert T == 7 and kinds == [0, 1, 2, 0, 3, 1] and model.p == 1
    fns = CModelFns(model)
    rng = np.random.default_rng(0)
    for t in range(6):
        x, u = rng.standard_normal(2), rng.standard_normal(1)
        y, fx, fu = fns.dyn(x, u, np.array([float(kinds[t])]))
        want = np.zeros(2); dyn[t].evaluate(want, x, u, None)
        np.testing.assert_allclose(y, want, rtol=1e-14)
        wfx = np.zeros((2, 2)); dyn[t].jacobian_state(wfx, x, u, None)
        np.testing.assert_allclose(np.asarray(fx).reshape(2, 2, order="F"), wfx, rtol=1e-14)
        g = fns.cost(False, x, u, np.array([float(kinds[t])]))[0]
        wg = np.zeros(1); obj[t].evaluate(wg, x, u, None)
        np.testing.assert_allclose(g, wg[0], rtol=1e-14)


def test_time_varying_dimensions_are_embedded_in_the_largest():
    """num_state / num_action / num_next_state differing between steps (src/dynamics.jl:5, src/solver.jl:28-30): the merged
    model works on the largest dimensions; every variant must evaluate like its own object on the leading components,
    produce zeros behind them, and charge u_a^2 / 2 for the action components it does not have."""
    from ilqr_b200 import Cost, Dynamics, dot
    from ilqr_b200.solver import _model_from_lists
    from oracle.c_oracle import CModelFns
    d1 = Dynamics(lambda x, u: [x[0] + 0.1 * x[1], x[1] + 0.1 * u[0], x[0] * u[0]], 2, 1)                         # 2 -> 3 states
    d2 = Dynamics(lambda x, u: [x[0] + 0.1 * x[1] + 0.05 * x[2], 0.9 * x[1] + 0.1 * u[0] - 0.05 * u[1]], 3, 2)   # 3 -> 2 states
    c1 = Cost(lambda x, u: dot(x, x) + 0.1 * dot(u, u), 2, 1)
    c2 = Cost(lambda x, u: dot(x, x) + 0.1 * dot(u, u) + 0.3 * x[2] * u[1], 3, 2)
    cT = Cost(lambda x, u: 10 * dot(x, x), 2, 0)
    dyn, obj = [d1, d2, d1, d2], [c1, c2, c1, c2, cT]
    model, T, kinds, dims = _model_from_lists(dyn, obj, None)
    assert (T, kinds, model.n, model.m, model.p) == (5, [0, 1, 0, 1], 3, 2, 1)
    assert dims == ([2, 3, 2, 3, 2], [1, 2, 1, 2])
    fns = CModelFns(model)
    rng = np.random.default_rng(1)
    for t in range(4):
        n_t, m_t, n_next = dims[0][t], dims[1][t], dims[0][t + 1]
        x, u = np.zeros(3), np.zeros(2)
        x[:n_t], u[:m_t] = rng.standard_normal(n_t), rng.standard_normal(m_t)
        y, fx, fu = fns.dyn(x, u, np.array([float(kinds[t])]))
        want = np.zeros(n_next); dyn[t].evaluate(want, x[:n_t], u[:m_t], None)
        np.testing.assert_allclose(y[:n_next], want, rtol=1e-14)
        assert not np.any(y[n_next:])
        wfx = np.zeros((n_next, n_t)); dyn[t].jacobian_state(wfx, x[:n_t], u[:m_t], None)
        fx = np.asarray(fx).reshape(3, 3, order="F")
        np.testing.assert_allclose(fx[:n_next, :n_t], wfx, rtol=1e-14)
        assert not np.any(fx[n_next:]) and not np.any(fx[:, n_t:])
        u[m_t:] = 0.7  # a value the solver never produces there: the padding term alone
        g = fns.cost(False, x, u, np.array([float(kinds[t])]))[0]
        wg = np.zeros(1); obj[t].evaluate(wg, x[:n_t], u[:m_t], None)
        np.testing.assert_allclose(g, wg[0] + 0.5 * 0.49 * (2 - m_t), rtol=1e-14)
    with pytest.raises(AssertionError):  # a chain whose state sizes do not connect (src/data/problem.jl shapes)
        _model_from_lists([d1, d1], [c1, c1, cT], None)
    # the embedded solve reproduces the solve on the true per-step shapes (literal oracle, as the reference works)
    from oracle.c_oracle import COracle
    from oracle.ilqr_oracle import Options as PyOptions, OracleSolver, rollout as py_rollout
    x1 = np.array([1.0, -0.5])
    ubar = [0.3 * rng.standard_normal(d.num_action) for d in dyn]
    xbar = py_rollout(dyn, x1, ubar)
    po = OracleSolver(dyn, obj, None, options=PyOptions(verbose=False))
    po.initialize_controls(list(ubar)); po.initialize_states(xbar); po.solve()
    xo, uo = po.get_trajectory()
    co = COracle(model, T, 1)
    w = np.zeros((1, T, 1)); w[0, : T - 1, 0] = kinds
    up, xp = np.zeros((1, T - 1, 2)), np.zeros((1, T, 3))
    for t in range(T - 1):
        up[0, t, : dims[1][t]] = ubar[t]
    for t in range(T):
        xp[0, t, : dims[0][t]] = xbar[t]
    co.set_parameters(w); co.initialize_controls(up); co.initialize_states(xp); co.solve()
    xc, uc = co.get_trajectory()
    assert int(co.get_stats()["iterations"][0]) == po.iterations[0]
    for t in range(T):
        np.testing.assert_allclose(xc[0, t, : dims[0][t]], xo[t], rtol=0, atol=1e-12)
        assert not np.any(xc[0, t, dims[0][t]:])
    for t in range(T - 1):
        np.testing.assert_allclose(uc[0, t, : dims[1][t]], uo[t], rtol=0, atol=1e-12)
        assert not np.any(uc[0, t, dims[1][t]:])


def test_explicit_derivative_constructors_take_c_snippets():
    """Dynamics(f, fx, fu, ...) / Constraint(f, fx, fu, ...) with user-provided functions (src/dynamics.jl:55-60,
    src/constraints.jl:54-64): here the three functions are C statement bodies compiled as they are.  The particle model
    given that way must evaluate exactly like the traced one (through the emitted C, host build)."""
    from ilqr_b200 import Cost, constraint_from_c, dot, dynamics_from_c
    from ilqr_b200.api import Model
    from oracle.c_oracle import CModelFns
    dyn = dynamics_from_c("y[0] = x[0] + x[1];\ny[1] = x[1] + u[0];",
                          "fx[0] = 1.0; fx[1] = 0.0; fx[2] = 1.0; fx[3] = 1.0;",
                          "fu[0] = 0.0; fu[1] = 1.0;", 2, 1)
    goal = constraint_from_c("c[0] = x[0] - 1.0; c[1] = x[1];", "cx[0] = 1.0; cx[1] = 0.0; cx[2] = 0.0; cx[3] = 1.0;", "", 2, 2, 0)
    stage = Cost(lambda x, u: 0.1 * dot(x, x) + 0.1 * dot(u, u), 2, 1)
    term = Cost(lambda x, u: 0.1 * dot(x, x), 2, 0)
    m = Model("particle_c", dyn, stage, term, None, goal)
    ref = problems.particle()
    assert (m.n, m.m, m.cs, m.ct) == (ref.n, ref.m, ref.cs, ref.ct)
    a, b = CModelFns(m), CModelFns(ref)
    rng = np.random.default_rng(1)
    for _ in range(5):
        x, u = rng.standard_normal(2), rng.standard_normal(1)
        for got, want in zip(a.dyn(x, u), b.dyn(x, u)):
            np.testing.assert_array_equal(got, want)
        for got, want in zip(a.con(True, x), b.con(True, x)):
            if got is not None:
                np.testing.assert_array_equal(got, want)
    with pytest.raises(NotImplementedError):
        dyn.evaluate(np.zeros(2), np.zeros(2), np.zeros(1), None)


def test_shard_bounds_cover_batch():
    from ilqr_b200.distributed import shard_bounds
    for B in (1, 7, 4096, 16384, 10):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_prints_contract_line():
    """bench.py --impl reference needs no GPU: it times the CPU restatement and prints one JSON line."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["metric"] == "ilqr_solves_per_sec_batched_acrobot_T101" and line["higher_is_better"] is True
