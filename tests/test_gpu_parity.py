"""PARITY (GPU): the CUDA engine, called through the C ABI (libilqr_cuda.so), against the CPU
oracle on the same seeded inputs.  North-star tolerance: identical iteration counts, cost /
violation history to 1e-9 relative, final trajectories to 1e-7.  The engine and the C oracle
implement one arithmetic contract, so the tests additionally demand BIT equality."""
import os

import numpy as np
import pytest

import ilqr_b200
from common import assert_same_solution, collect, inputs, lq_inputs
from ilqr_b200 import build, capi, problems
from oracle.c_oracle import COptions, COracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_pair(name, B, T, seed=0, cap=1000, variant="", **opts):
    model, x1, ubar = inputs(name, B, T, seed)
    co = COracle(model, T, B, options=COptions.default(**opts), history_cap=cap)
    go = capi.default_options()
    for k, v in opts.items():
        setattr(go, k, {"armijo": 0, "none": 1}[v] if k == "line_search" else v)
    h = capi.Handle(build.model_library(model, variant=variant), T, model.n, model.m, model.p, model.cs, model.ct, B,
                    history_cap=cap, options=go)
    return model, x1, ubar, co, h


def force_tp(monkeypatch, back=True, fwd=True, compact=None, tma=None):
    """Engine tuning knobs read at ilqr_create: which grids take the thread-per-problem kernels, and down to how many
    32-problem blocks a streamed job's grid is compacted in its drain phase."""
    never = str(1 << 40)
    monkeypatch.setenv("ILQR_TP_MIN_BLOCKS", "0" if back else never)
    monkeypatch.setenv("ILQR_FT_MIN_BLOCKS", "0" if fwd else never)
    if compact is not None:
        monkeypatch.setenv("ILQR_COMPACT_MIN_BLOCKS", str(compact))
    if tma is not None:
        monkeypatch.setenv("ILQR_FWD_TMA", str(int(tma)))


def solve_both(co, h, x1, ubar):
    xo = co.rollout(x1, ubar)
    xg = h.rollout(x1, ubar)
    np.testing.assert_array_equal(xg, xo)  # rollout (src/rollout.jl:33-42)
    for s in (co, h):
        s.initialize_controls(ubar)
        s.initialize_states(xo)
        s.solve()
    return xo


@pytest.mark.parametrize("name,T,B", [("particle", 11, 40), ("car", 51, 70), ("acrobot", 51, 64), ("pendulum", 31, 33),
                                      ("acrobot", 101, 96)])
def test_solve_matches_oracle(name, T, B):
    model, x1, ubar, co, h = make_pair(name, B, T, seed=1)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))
    if model.constrained:
        dg, do = h.get_duals(), co.get_duals()
        for k in ("dual", "penalty", "violations", "active_set"):
            np.testing.assert_array_equal(dg[k], do[k], err_msg=k)
    Kg, kg = h.get_policy()
    Ko, ko = co.get_policy()
    np.testing.assert_array_equal(Kg, Ko)
    np.testing.assert_array_equal(kg, ko)
    xg, ug = h.get_trajectory(current=True)
    xo, uo = co.get_trajectory(current=True)
    np.testing.assert_array_equal(xg, xo)
    np.testing.assert_array_equal(ug, uo)


@pytest.mark.parametrize("name,T,B", [("particle", 11, 40), ("car", 31, 70), ("acrobot", 51, 64), ("pendulum", 31, 33)])
@pytest.mark.parametrize("how", ["tp", "tpback", "tpfwd", "nohacc", "nohacc-tp", "tma", "notma", "nohacc-tma", "sring", "nohacc-sring"])
def test_kernel_variants(name, T, B, how, monkeypatch):
    """The other decompositions of the gradients! + backward_pass! tick must give the oracle's bits too:
    "tp"     -- k_forward_tp + k_linback_tp (one thread per problem, no warp specialisation), which the engine picks by
                itself only for dense grids (ILQR_TP_MIN_BLOCKS / ILQR_FT_MIN_BLOCKS = 0 force them; "tpback" / "tpfwd":
                only one of the two);
    "tma"    -- k_forward_tma (one TMA-filled shared-memory ring per CTA feeding both trial warps and the
                expected-decrease warp) against "notma" = k_forward (a private cp.async ring per warp); "sring" = the same
                shared ring fed by the producer warp's cp.async copies instead of bulk copies;
    "nohacc" -- per-time-step Hessian accumulators (-DILQR_NO_HACC) on models whose constant stage Hessians would
                otherwise take the one-accumulator-per-problem shortcut (HACC in csrc/ilqr_kernels.cuh)."""
    force_tp(monkeypatch, back="tp" in how and how != "tpfwd", fwd="tp" in how and how != "tpback",
             tma=1 if how.endswith("tma") and how != "notma" else (2 if how.endswith("sring") else (0 if how == "notma" else None)))
    model, x1, ubar, co, h = make_pair(name, B, T, seed=21, variant="nohacc" if "nohacc" in how else "")
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))
    Kg, kg = h.get_policy(); Ko, ko = co.get_policy()
    np.testing.assert_array_equal(Kg, Ko)
    np.testing.assert_array_equal(kg, ko)
    ao, xo = co.mpc_step(); ag, xg = h.mpc_step()   # a warm re-solve on the same handle (accumulators restart)
    np.testing.assert_array_equal(ag, ao)
    assert_same_solution(collect(h), collect(co))


@pytest.mark.parametrize("name", ["particle", "car", "acrobot", "pendulum"])
def test_golden_fixtures(name):
    """Committed fixtures generated by tests/golden/make_golden.py."""
    from make_golden import CASES
    T, B, seed = CASES[name]
    g = np.load(os.path.join(GOLD, f"{name}_c.npz"))
    model, x1, ubar = inputs(name, B, T, seed)
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B)
    xbar = h.rollout(x1, ubar)
    np.testing.assert_array_equal(xbar, g["xbar"])
    h.initialize_controls(ubar); h.initialize_states(xbar); h.solve()
    st, hist = h.get_stats(), h.get_history()
    np.testing.assert_array_equal(st["iterations"], g["iterations"])
    n = g["cost"].shape[1]
    np.testing.assert_allclose(hist["cost"][:, :n], g["cost"], rtol=1e-9, atol=0)
    np.testing.assert_allclose(hist["max_violation"][:, :n], g["viol"], rtol=1e-9, atol=0)
    x, u = h.get_trajectory()
    np.testing.assert_allclose(x, g["x"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(u, g["u"], rtol=0, atol=1e-7)
    np.testing.assert_array_equal(x, g["x"])
    # and the literal numpy/LAPACK oracle's fixture, at the north-star tolerance on the trajectory
    gp = np.load(os.path.join(GOLD, f"{name}_py.npz"))
    np.testing.assert_array_equal(st["iterations"], gp["iterations"])
    np.testing.assert_allclose(x, gp["x"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("name,T,B", [("car", 31, 40), ("acrobot", 31, 33), ("pendulum", 21, 20), ("lq8", 21, 24)])
def test_large_model_kernels_on_small_models(name, T, B):
    """The loop-based kernels meant for wide models (csrc/ilqr_large_*.cuh: CTA-per-problem Riccati with the
    value function in shared memory, streamed rollout / linearisation) compiled for small models
    (-DILQR_FORCE_LARGE) must reproduce the oracle bit for bit too -- including AL terms and parameters."""
    if name == "lq8":
        model, x1, ubar, w = lq_inputs(B, T, seed=5)
        kw = dict(objective_tolerance=0.0, lagrangian_gradient_tolerance=0.0, max_iterations=3)
    else:
        model, x1, ubar = inputs(name, B, T, seed=5)
        w, kw = None, {}
    co = COracle(model, T, B, options=COptions.default(**kw))
    go = capi.default_options()
    for k, v in kw.items():
        setattr(go, k, v)
    h = capi.Handle(build.model_library(model, variant="large"), T, model.n, model.m, model.p, model.cs, model.ct, B, options=go)
    if w is not None:
        co.set_parameters(w); h.set_parameters(w)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))
    Kg, kg = h.get_policy(); Ko, ko = co.get_policy()
    np.testing.assert_array_equal(Kg, Ko)
    np.testing.assert_array_equal(kg, ko)
    ao, xo = co.mpc_step(); ag, xg = h.mpc_step()
    np.testing.assert_array_equal(ag, ao)
    assert_same_solution(collect(h), collect(co))


def test_wide_model_n64_matches_oracle():
    """BASELINE config 4's shapes (n = 64, m = 16, p = 128) on the wide-model path: CTA-per-problem Riccati kernel
    with the value function in shared memory, against the oracle, bit for bit.  (Banded plant so that the generated
    model compiles in about a minute; the dense plant of benchmarks/c4.py takes ~35 minutes of nvcc time.)"""
    B, T = 12, 24
    model = problems.lq_banded(64, 16)
    rng = np.random.default_rng(9)
    w = np.zeros((B, T, 128)); w[:, :, :64] = rng.uniform(-1, 1, (B, 1, 64))
    w[:, :, 64:] = np.sin(0.3 * np.arange(T)[None, :, None] + rng.uniform(0, 6, (B, 1, 64)))
    x1 = rng.standard_normal((B, 64)); ubar = 0.2 * rng.standard_normal((B, T - 1, 16))
    kw = dict(objective_tolerance=0.0, lagrangian_gradient_tolerance=0.0, max_iterations=3)
    co = COracle(model, T, B, options=COptions.default(**kw))
    go = capi.default_options()
    for k, v in kw.items():
        setattr(go, k, v)
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B, options=go)
    co.set_parameters(w); h.set_parameters(w)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))
    Kg, kg = h.get_policy(); Ko, ko = co.get_policy()
    np.testing.assert_array_equal(Kg, Ko)
    np.testing.assert_array_equal(kg, ko)
    assert collect(h)["stats"]["flags"].max() == 0


def test_parameters_and_unfused_path():
    """A model with per-step parameters (src/data/problem.jl:25-30) that is too wide for the fused kernel's
    shared-memory ring: exercises ilqr_set_parameters and the k_linearize + k_backward pair."""
    import torch
    B, T = 40, 21
    model, x1, ubar, w = lq_inputs(B, T, seed=3)
    co = COracle(model, T, B, options=COptions.default(objective_tolerance=0.0, lagrangian_gradient_tolerance=0.0, max_iterations=4))
    go = capi.default_options(); go.objective_tolerance = 0.0; go.lagrangian_gradient_tolerance = 0.0; go.max_iterations = 4
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B, options=go)
    co.set_parameters(w); h.set_parameters(w)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))
    assert collect(h)["stats"]["iterations"].max() >= 2
    # streamed, with the parameters travelling with each problem
    xbar = co.rollout(x1, ubar)
    co2 = COracle(model, T, B, options=co.options); co2.set_parameters(w)
    co2.initialize_controls(ubar); co2.initialize_states(xbar); co2.solve()
    h2 = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, 16, options=go)
    dx, du, dw = torch.from_numpy(xbar).cuda(), torch.from_numpy(ubar).cuda(), torch.from_numpy(w).cuda()
    ox, ou = torch.zeros_like(dx), torch.zeros_like(du)
    h2.solve_stream(B, dx.data_ptr(), du.data_ptr(), dw.data_ptr(), ox.data_ptr(), ou.data_ptr())
    torch.cuda.synchronize()
    xo, uo = co2.get_trajectory()
    np.testing.assert_array_equal(ox.cpu().numpy(), xo)
    np.testing.assert_array_equal(ou.cpu().numpy(), uo)


@pytest.mark.parametrize("B", [1, 31, 32, 33, 100])
def test_ragged_batch_sizes(B):
    """batch not a multiple of the warp size: padded lanes must stay inert."""
    model, x1, ubar, co, h = make_pair("particle", B, 11, seed=B)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))


def test_minimal_horizon():
    model, x1, ubar, co, h = make_pair("particle", 5, 2, seed=2)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))


@pytest.mark.parametrize("opts", [
    dict(line_search="none", max_iterations=6, max_dual_updates=3),      # Q12
    dict(max_iterations=3, max_dual_updates=2),                          # iteration limits bind
    dict(reset_cache=1, max_dual_updates=3),                             # src/solve.jl:12
    dict(min_step_size=0.3),                                             # 2 line-search trials only
    dict(min_step_size=2.0, max_dual_updates=2),                         # zero trials
    dict(max_iterations=0, max_dual_updates=2),
    dict(max_dual_updates=0),
    dict(min_step_size=1e-9),                                            # trial cap of 25 binds (src/forward_pass.jl:29)
    dict(initial_constraint_penalty=50.0, scaling_penalty=3.0, max_penalty=200.0, constraint_tolerance=1e-4),
    dict(objective_tolerance=1e-7, lagrangian_gradient_tolerance=1e-7, max_iterations=30),
])
@pytest.mark.parametrize("kern", ["default", "tp", "tma"])
def test_options(opts, kern, monkeypatch):
    if kern == "tp":
        force_tp(monkeypatch)
    elif kern == "tma":
        monkeypatch.setenv("ILQR_FWD_TMA", "1")
    model, x1, ubar, co, h = make_pair("car", 37, 31, seed=4, **opts)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))


def test_warm_start_resolves_keep_solver_state():
    """solve!(solver, x, u) twice on one solver (src/solve.jl:131-135): the current trajectory
    and stale constraint values persist between solves (Q2), the engine must carry them too."""
    model, x1, ubar, co, h = make_pair("car", 40, 31, seed=6)
    xbar = solve_both(co, h, x1, ubar)
    first = collect(h)
    for rep in range(2):
        co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
        h.solve_warm(xbar, ubar)
        assert_same_solution(collect(h), collect(co))
    # the warm re-solve is NOT a replay of the first solve (current trajectory differs)
    assert not np.array_equal(first["history"]["cost"], collect(h)["history"]["cost"])


def test_set_options_between_solves():
    model, x1, ubar, co, h = make_pair("particle", 8, 11, seed=8)
    solve_both(co, h, x1, ubar)
    o = COptions.default(max_iterations=2, max_dual_updates=2)
    co.set_options(o)
    go = capi.default_options(); go.max_iterations = 2; go.max_dual_updates = 2
    h.set_options(go)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))


def test_mpc_steps_match_oracle():
    """receding-horizon closed loop (BASELINE config 5) for a few steps."""
    model, x1, ubar, co, h = make_pair("acrobot", 48, 31, seed=9)
    solve_both(co, h, x1, ubar)
    for step in range(4):
        ao, xo = co.mpc_step()
        ag, xg = h.mpc_step()
        np.testing.assert_array_equal(ag, ao)
        np.testing.assert_array_equal(xg, xo)
        assert_same_solution(collect(h), collect(co))


@pytest.mark.parametrize("name,T,B,steps", [("acrobot", 31, 48, 12), ("car", 21, 40, 8), ("pendulum", 21, 33, 10)])
def test_asynchronous_mpc_run_matches_oracle(name, T, B, steps):
    """ilqr_mpc_run: every problem does `steps` receding-horizon re-solves at its own pace; the applied
    actions, plant states and final trajectories must equal the oracle's lock-step closed loop."""
    import torch
    model, x1, ubar, co, h = make_pair(name, B, T, seed=15)
    solve_both(co, h, x1, ubar)
    au = torch.zeros((steps, B, model.m), dtype=torch.float64, device="cuda")
    xn = torch.zeros((steps, B, model.n), dtype=torch.float64, device="cuda")
    tot = torch.zeros(B, dtype=torch.int32, device="cuda")
    h.mpc_run(steps, au.data_ptr(), xn.data_ptr(), tot.data_ptr())
    torch.cuda.synchronize()
    total = np.zeros(B, np.int64)
    prev = co.get_stats()["iterations"].astype(np.int64)
    for s in range(steps):
        ao, xo = co.mpc_step()
        # iterations of THIS re-solve: data.iterations is cumulative for unconstrained solvers (src/solve.jl:137-139
        # never resets it), reset by every constrained solve (src/solve.jl:93)
        now = co.get_stats()["iterations"].astype(np.int64)
        total += now if model.constrained else now - prev
        prev = now
        np.testing.assert_array_equal(au[s].cpu().numpy(), ao, err_msg=f"applied action, step {s}")
        np.testing.assert_array_equal(xn[s].cpu().numpy(), xo, err_msg=f"plant state, step {s}")
    np.testing.assert_array_equal(tot.cpu().numpy(), total)
    got, ref = collect(h), collect(co)
    np.testing.assert_array_equal(got["x"], ref["x"])
    np.testing.assert_array_equal(got["u"], ref["u"])
    np.testing.assert_array_equal(got["stats"]["iterations"], ref["stats"]["iterations"])
    # and the synchronous single-step entry still agrees afterwards
    ao, xo = co.mpc_step()
    ag, xg = h.mpc_step()
    np.testing.assert_array_equal(ag, ao)
    assert_same_solution(collect(h), collect(co))


def test_history_cap_truncates_without_changing_results():
    model, x1, ubar, co, h = make_pair("car", 16, 31, seed=10, cap=5)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h, 5), collect(co, 5))


def test_failure_flags_and_error_paths():
    model, x1, ubar, co, h = make_pair("particle", 4, 11, seed=11)
    with pytest.raises(ValueError):
        h.initialize_controls(ubar[:, :-1])
    with pytest.raises(capi.IlqrError):
        h.get_history(cap=5000)
    bad = capi.default_options(); bad.line_search = 7
    with pytest.raises(capi.IlqrError, match="line_search"):
        h.set_options(bad)
    # non-finite input: the engine must terminate and flag, like the oracle
    u2 = ubar.copy(); u2[0, 3, 0] = np.nan
    xb = co.rollout(x1, ubar)
    for s in (co, h):
        s.initialize_controls(u2); s.initialize_states(xb); s.solve()
    sg, so = h.get_stats(), co.get_stats()
    np.testing.assert_array_equal(sg["iterations"], so["iterations"])
    np.testing.assert_array_equal(sg["flags"], so["flags"])
    assert sg["flags"][0] & 2 and not sg["flags"][1]


def test_device_pointer_entry_points():
    import torch
    model, x1, ubar, co, h = make_pair("acrobot", 64, 51, seed=12)
    xbar = solve_both(co, h, x1, ubar)
    ref = collect(h)
    h2 = capi.Handle(build.model_library(model), 51, model.n, model.m, model.p, model.cs, model.ct, 64)
    dx, du = torch.from_numpy(xbar).cuda(), torch.from_numpy(ubar).cuda()
    stream = torch.cuda.Stream()
    h2.set_stream(stream.cuda_stream)
    h2.initialize_controls_device(du.data_ptr()); h2.initialize_states_device(dx.data_ptr()); h2.solve()
    ox = torch.empty_like(dx); ou = torch.empty_like(du)
    h2.get_trajectory_device(ox.data_ptr(), ou.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ox.cpu().numpy(), ref["x"])
    np.testing.assert_array_equal(ou.cpu().numpy(), ref["u"])
    c = h2.get_counters()
    assert c["ticks"] > 0 and c["launches"] >= 2 * (c["ticks"] - 2)


@pytest.mark.parametrize("name,T,slots,n", [("acrobot", 51, 64, 300), ("car", 31, 32, 150), ("pendulum", 31, 96, 50),
                                             ("particle", 11, 33, 200), ("acrobot", 31, 512, 1500), ("car", 21, 300, 700)])
@pytest.mark.parametrize("kern", ["default", "tp", "default-compact", "tp-compact", "tma-compact", "sring-compact"])
def test_streaming_matches_fresh_oracle_solves(name, T, slots, n, kern, monkeypatch):
    """ilqr_solve_stream (continuous batching): n problems through `slots` slots; every problem must come out
    exactly as a fresh solver would solve it, whatever slot / tick it ran in -- with either kernel set, and with the
    drain compaction (running problems packed into the lowest slots, smaller grids) allowed down to one block."""
    import torch
    force_tp(monkeypatch, back="tp" in kern, fwd="tp" in kern, compact=1 if "compact" in kern else 1 << 40,
             tma=1 if "tma" in kern else (2 if "sring" in kern else None))
    model, x1, ubar = inputs(name, n, T, seed=31)
    co = COracle(model, T, n, history_cap=1)
    xbar = co.rollout(x1, ubar)
    co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
    xo, uo = co.get_trajectory()
    so = co.get_stats()
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, slots, history_cap=1)
    dx, du = torch.from_numpy(xbar).cuda(), torch.from_numpy(ubar).cuda()
    ox, ou = torch.zeros_like(dx), torch.zeros_like(du)
    it = torch.zeros(n, dtype=torch.int32, device="cuda"); st = torch.zeros(n, dtype=torch.uint8, device="cuda")
    J = torch.zeros(n, dtype=torch.float64, device="cuda"); mv = torch.zeros_like(J); ss = torch.zeros_like(J)
    fl = torch.zeros(n, dtype=torch.int32, device="cuda")
    for rep in range(2):  # a second job on the same handle must not see leftovers of the first
        ox.zero_(); ou.zero_(); it.zero_()
        h.solve_stream(n, dx.data_ptr(), du.data_ptr(), 0, ox.data_ptr(), ou.data_ptr(), it.data_ptr(), st.data_ptr(),
                       J.data_ptr(), mv.data_ptr(), ss.data_ptr(), fl.data_ptr())
        torch.cuda.synchronize()
        np.testing.assert_array_equal(it.cpu().numpy(), so["iterations"])
        np.testing.assert_array_equal(ox.cpu().numpy(), xo)
        np.testing.assert_array_equal(ou.cpu().numpy(), uo)
        np.testing.assert_array_equal(J.cpu().numpy(), so["objective"])
        np.testing.assert_array_equal(mv.cpu().numpy(), so["max_violation"])
        np.testing.assert_array_equal(ss.cpu().numpy(), so["step_size"])
        np.testing.assert_array_equal(st.cpu().numpy(), so["status"])
        np.testing.assert_array_equal(fl.cpu().numpy().astype(np.uint32), so["flags"])
    if "compact" in kern and slots >= 64 and n > slots:
        assert h.get_counters()["compactions"] >= 2, "the drain phase of both jobs should have packed the slots"
    elif "compact" not in kern:
        assert h.get_counters()["compactions"] == 0
    # the handle still works in batch mode afterwards
    if slots <= n:
        h.initialize_controls(ubar[:slots]); h.initialize_states(xbar[:slots]); h.solve()
        assert h.get_stats()["iterations"].max() > 0


def test_host_buffer_streaming_api():
    """ilqr_solve_stream_host through the Python API: host arrays in, host arrays out."""
    from ilqr_b200 import Options, Solver, solve_stream
    n, T, slots = 130, 31, 32
    model, x1, ubar = inputs("car", n, T, seed=41)
    co = COracle(model, T, n, history_cap=1)
    xbar = co.rollout(x1, ubar)
    co.initialize_controls(ubar); co.initialize_states(xbar); co.solve()
    xo, uo = co.get_trajectory()
    solver = Solver(model, T=T, batch=slots, options=Options(verbose=False))
    x, u, st = solve_stream(solver, xbar, ubar)
    np.testing.assert_array_equal(x, xo)
    np.testing.assert_array_equal(u, uo)
    np.testing.assert_array_equal(st["iterations"], co.get_stats()["iterations"])
    np.testing.assert_array_equal(st["max_violation"], co.get_stats()["max_violation"])


def test_reference_facing_python_api():
    """The user-level calls with the reference's names (src/IterativeLQR.jl:31-45), batch = 1,
    on the README quick start (examples/particle.jl)."""
    from ilqr_b200 import (Constraint, Cost, Dynamics, Options, Solver, dot, get_trajectory, initialize_controls,
                           initialize_states, rollout, solve)
    T, n, m = 11, 2, 1
    particle = Dynamics(problems.particle_discrete, n, m)
    dynamics = [particle for _ in range(T - 1)]
    rng = np.random.default_rng(0)
    x1, xT = np.zeros(2), np.array([1.0, 0.0])
    ubar = [1.0e-1 * rng.standard_normal(m) for _ in range(T - 1)]
    xbar = rollout(dynamics, x1, ubar)
    stage = Cost(lambda x, u: 0.1 * dot(x, x) + 0.1 * dot(u, u), n, m)
    term = Cost(lambda x, u: 0.1 * dot(x, x), n, 0)
    objective = [stage for _ in range(T - 1)] + [term]
    no_con = Constraint()
    constraints = [no_con for _ in range(T - 1)] + [Constraint(lambda x, u: x - xT, n, 0)]
    solver = Solver(dynamics, objective, constraints, options=Options(verbose=False))
    initialize_controls(solver, ubar)
    initialize_states(solver, xbar)
    solve(solver)
    x_sol, u_sol = get_trajectory(solver)
    assert len(x_sol) == T and len(u_sol) == T - 1
    assert np.max(np.abs(x_sol[-1] - xT)) < 5e-3
    # same thing through the oracle
    co = COracle(solver.model, T, 1)
    co.initialize_controls(np.array(ubar)[None]); co.initialize_states(np.array(xbar)[None]); co.solve()
    xo, uo = co.get_trajectory()
    np.testing.assert_array_equal(np.array(x_sol), xo[0])
    np.testing.assert_array_equal(np.array(u_sol), uo[0])


@pytest.mark.parametrize("name,T,B", [("acrobot", 101, 4096), ("car", 51, 2048)])
def test_full_size_batches(name, T, B):
    """BASELINE.json sizes (configs[1]; configs[2]'s per-GPU shard): every problem bit-equal to the
    oracle, plus size-independent properties: feasibility at the reference's tolerance, idempotent
    re-solve of the converged solution, and shard-independence (a problem's result does not depend
    on which batch it is solved in)."""
    model, x1, ubar, co, h = make_pair(name, B, T, seed=100, cap=8)
    xbar = solve_both(co, h, x1, ubar)
    got, ref = collect(h, 8), collect(co, 8)
    assert_same_solution(got, ref)
    st = got["stats"]
    assert np.mean(st["max_violation"] <= 5e-3) > 0.99
    # shard independence: solve problems [100, 164) alone
    sl = slice(100, 164)
    h2 = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, 64, history_cap=8)
    h2.initialize_controls(ubar[sl]); h2.initialize_states(xbar[sl]); h2.solve()
    x2, u2 = h2.get_trajectory()
    np.testing.assert_array_equal(x2, got["x"][sl])
    np.testing.assert_array_equal(h2.get_stats()["iterations"], st["iterations"][sl])


@pytest.mark.parametrize("name,T,B", [("acrobot", 21, 9728), ("car", 11, 9600)])
@pytest.mark.parametrize("ring", ["shared", "private"])
def test_dense_grid_uses_the_register_capped_forward_kernel(name, T, B, ring, monkeypatch):
    """Grids beyond two CTAs per SM (B > 9472 on a 148-SM B200) run the register-capped instantiations (three CTAs per
    SM, 168 registers): by default k_forward_tma<3, false> (one cp.async-fed ring per CTA), with ILQR_FWD_TMA=0
    k_forward<3> (a private ring per warp); the dense k_linback as well.  Results must be the oracle's bit for bit."""
    if ring == "private":
        monkeypatch.setenv("ILQR_FWD_TMA", "0")
    model, x1, ubar, co, h = make_pair(name, B, T, seed=7, cap=8)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h, 8), collect(co, 8))


def test_time_varying_stage_functions():
    """Per-step distinct Dynamics / Cost objects through the reference-facing API (src/solver.jl:28-30): the engine
    solves the merged model (variant selected by a trailing parameter); the literal oracle solves the SAME problem from
    the per-step objects themselves (lambdified, no merging) -- north-star tolerance -- and the C oracle, compiled from
    the merged header, must agree bit for bit."""
    from ilqr_b200 import Constraint, Cost, Dynamics, Options, Solver, dot, get_trajectory, initialize_controls, initialize_states, rollout, solve
    from oracle.ilqr_oracle import Options as PyOptions, OracleSolver, rollout as py_rollout
    n, m, T = 2, 1, 9
    d1 = Dynamics(lambda x, u: [x[0] + 0.1 * x[1], x[1] + 0.1 * u[0]], n, m)
    d2 = Dynamics(lambda x, u: [x[0] + 0.1 * x[1], 0.9 * x[1] + 0.1 * u[0] - 0.05 * x[0] ** 3], n, m)
    c1 = Cost(lambda x, u: dot(x, x) + 0.1 * dot(u, u), n, m)
    c2 = Cost(lambda x, u: 2 * dot(x, x) + 0.1 * dot(u, u) + 0.01 * u[0] ** 4, n, m)
    cT = Cost(lambda x, u: 10 * dot(x, x), n, 0)
    goal = Constraint(lambda x, u: [x[0] - 0.5, x[1]], n, 0)
    none = Constraint()
    dyn = [d1, d2, d1, d1, d2, d2, d1, d2]
    obj = [c1, c1, c2, c1, c2, c1, c2, c2, cT]
    con = [none] * (T - 1) + [goal]
    rng = np.random.default_rng(3)
    x1 = np.array([1.0, -0.5])
    ubar = [0.3 * rng.standard_normal(m) for _ in range(T - 1)]
    xbar = rollout(dyn, x1, ubar)
    np.testing.assert_allclose(np.array(xbar), np.array(py_rollout(dyn, x1, ubar)), rtol=0, atol=1e-13)
    solver = Solver(dyn, obj, con, options=Options(verbose=False))
    initialize_controls(solver, ubar); initialize_states(solver, xbar); solve(solver)
    x, u = get_trajectory(solver)
    s = OracleSolver(dyn, obj, con, options=PyOptions(verbose=False))
    s.initialize_controls(list(ubar)); s.initialize_states(xbar); s.solve()
    xo, uo = s.get_trajectory()
    assert int(solver.data["iterations"][0]) == s.iterations[0]
    np.testing.assert_allclose(np.array(x), np.array(xo), rtol=0, atol=1e-7)
    np.testing.assert_allclose(np.array(u), np.array(uo)[: T - 1], rtol=0, atol=1e-7)
    assert np.max(np.abs(np.array(x)[-1] - np.array([0.5, 0.0]))) < 5e-3
    co = COracle(solver.model, T, 1)
    w = np.zeros((1, T, 1)); w[0, : T - 1, 0] = solver.stage_kinds
    co.set_parameters(w)
    co.initialize_controls(np.array(ubar)[None]); co.initialize_states(np.array(xbar)[None]); co.solve()
    xc, uc = co.get_trajectory()
    np.testing.assert_array_equal(np.array(x), xc[0])
    np.testing.assert_array_equal(np.array(u), uc[0])


def test_time_varying_dimensions():
    """num_state / num_action / num_next_state differing between steps (src/dynamics.jl:5, src/solver.jl:28-30): the engine
    solves the model embedded in the largest dimensions (api.merge_stage_variants); the literal oracle works on the true
    per-step shapes, as the reference does.  Same iteration count, north-star tolerance on the trajectories, which come back
    in the steps' own lengths."""
    from ilqr_b200 import Constraint, Cost, Dynamics, Options, Solver, dot, get_trajectory, initialize_controls, initialize_states, solve
    from oracle.ilqr_oracle import Options as PyOptions, OracleSolver, rollout as py_rollout
    d1 = Dynamics(lambda x, u: [x[0] + 0.1 * x[1], x[1] + 0.1 * u[0], 0.5 * x[0] * u[0]], 2, 1)                            # 2 -> 3 states, 1 action
    d2 = Dynamics(lambda x, u: [x[0] + 0.1 * x[1] + 0.05 * x[2], 0.9 * x[1] + 0.1 * u[0] - 0.05 * u[1] - 0.05 * x[0] ** 3], 3, 2)  # 3 -> 2 states, 2 actions
    c1 = Cost(lambda x, u: dot(x, x) + 0.1 * dot(u, u), 2, 1)
    c2 = Cost(lambda x, u: dot(x, x) + 0.1 * dot(u, u) + 0.3 * x[2] * u[1], 3, 2)
    cT = Cost(lambda x, u: 10 * dot(x, x), 2, 0)
    goal = Constraint(lambda x, u: [x[0] - 0.5, x[1]], 2, 0)
    none = Constraint()
    T = 9
    dyn = [d1, d2] * 4
    obj = [c1, c2] * 4 + [cT]
    con = [none] * (T - 1) + [goal]
    rng = np.random.default_rng(5)
    x1 = np.array([1.0, -0.5])
    ubar = [0.3 * rng.standard_normal(d.num_action) for d in dyn]
    xbar = py_rollout(dyn, x1, ubar)
    assert [len(v) for v in xbar] == [2, 3] * 4 + [2]
    solver = Solver(dyn, obj, con, options=Options(verbose=False))
    assert (solver.model.n, solver.model.m) == (3, 2) and solver.dims == ([2, 3] * 4 + [2], [1, 2] * 4)
    initialize_controls(solver, ubar); initialize_states(solver, xbar); solve(solver)
    x, u = get_trajectory(solver)
    assert [len(v) for v in x] == [2, 3] * 4 + [2] and [len(v) for v in u] == [1, 2] * 4
    s = OracleSolver(dyn, obj, con, options=PyOptions(verbose=False))
    s.initialize_controls(list(ubar)); s.initialize_states(xbar); s.solve()
    xo, uo = s.get_trajectory()
    assert int(solver.data["iterations"][0]) == s.iterations[0] and s.iterations[0] > 3
    for t in range(T):
        np.testing.assert_allclose(x[t], xo[t], rtol=0, atol=1e-7)
    for t in range(T - 1):
        np.testing.assert_allclose(u[t], uo[t], rtol=0, atol=1e-7)
    assert np.max(np.abs(x[-1] - np.array([0.5, 0.0]))) < 5e-3
    # the padded components never leave zero
    xp, up = solver.handle.get_trajectory()
    assert not np.any(xp[0, 0::2, 2]) and not np.any(up[0, 0::2, 1])


def test_augmented_lagrangian_callback():
    """solve!(solver; augmented_lagrangian_callback!) (src/solve.jl:88,125) through ilqr_solve_outer: a callback that does
    nothing must leave the solve bit-identical to ilqr_solve on a batch; a callback that changes an option after every
    dual update must match the literal oracle running the same callback."""
    from ilqr_b200 import Options, Solver, get_trajectory, initialize_controls, initialize_states, solve
    from test_oracle import py_solver
    model, x1, ubar, co, h = make_pair("car", 24, 31, seed=13)
    xbar = solve_both(co, h, x1, ubar)
    ref = collect(h)
    calls = []
    solver = Solver(model, T=31, batch=24, options=Options(verbose=False))
    initialize_controls(solver, ubar); initialize_states(solver, xbar)
    solve(solver, augmented_lagrangian_callback=lambda s: calls.append(1))
    got = dict(stats=solver.handle.get_stats(), history=solver.handle.get_history(), x=solver.handle.get_trajectory()[0],
               u=solver.handle.get_trajectory()[1])
    assert_same_solution(got, ref)
    outer_max = int(ref["history"]["outer"].max())
    assert len(calls) in (outer_max - 1, outer_max)  # one call per round of dual updates of the batch
    # a callback with an effect, batch of one, against the literal oracle
    def cb_engine(s):  # from the second inner solve on: at most 3 iterations each, gentler penalty growth
        s.options.max_iterations = 3
        s.options.scaling_penalty = 3.0
    def cb_oracle(s):
        s.options.max_iterations = 3
        s.options.scaling_penalty = 3.0
    s1 = Solver(model, T=31, batch=1, options=Options(verbose=False))
    initialize_controls(s1, ubar[:1]); initialize_states(s1, xbar[:1])
    solve(s1, augmented_lagrangian_callback=cb_engine)
    x, u = get_trajectory(s1)
    po = py_solver(model, 31, x1[0], ubar[0])
    po.solve(augmented_lagrangian_callback=cb_oracle)
    xo, uo = po.get_trajectory()
    assert int(s1.data["iterations"][0]) == po.iterations[0]
    np.testing.assert_allclose(np.array(x), np.array(xo), rtol=0, atol=1e-7)
    np.testing.assert_allclose(np.array(u), np.array(uo)[:30], rtol=0, atol=1e-7)
    # ... and it did change the solve
    assert int(s1.data["iterations"][0]) != int(ref["stats"]["iterations"][0]) or not np.array_equal(np.array(x), ref["x"][0])


@pytest.mark.parametrize("variant", ["", "nodmma", "nojacconst", "nojacconst_nodmma", "nooverlap", "rl_fwarp", "rl_cholright", "rl_helpers"])
@pytest.mark.parametrize("wp", ["1", "0"])
def test_wide_dense_model_forward_kernels(wp, variant, monkeypatch):
    """BASELINE config 4's DENSE plant (n = 64, m = 16, p = 128; table-mode generated code) on the wide-model path, three
    iLQR iterations: k_forward_wp (a warp per problem and trial, matrix-vector outputs spread over the lanes) and the
    thread-per-problem forward kernel (ILQR_FWD_WP=0) must both reproduce the oracle bit for bit, Riccati kernel with the
    per-problem Hessian accumulator included -- with its dense contractions on the FP64 tensor cores (DMMA m8n8k4 tiles, the
    default) and as register-tiled DFMA loops (build variant "nodmma").  The plant is linear, so by default ONE staged Jacobian
    block serves the whole batch (JAC_CONST); the "nojacconst" variants keep the general path -- k_linearize writing one
    problem-major block per (problem, step) through its transposing tile, the Riccati kernel fetching them by bulk copy."""
    monkeypatch.setenv("ILQR_FWD_WP", wp)
    B, T = 6, 12
    model, x1, ubar, w = lq_inputs(B, T, 64, 16, seed=11)
    kw = dict(objective_tolerance=0.0, lagrangian_gradient_tolerance=0.0, max_iterations=3)
    co = COracle(model, T, B, options=COptions.default(**kw))
    go = capi.default_options()
    for k, v in kw.items():
        setattr(go, k, v)
    h = capi.Handle(build.model_library(model, variant=variant), T, model.n, model.m, model.p, model.cs, model.ct, B, options=go)
    co.set_parameters(w); h.set_parameters(w)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))
    assert collect(h)["stats"]["iterations"].min() >= 2 and collect(h)["stats"]["flags"].max() == 0
    ao, xo = co.mpc_step(); ag, xg = h.mpc_step()
    np.testing.assert_array_equal(ag, ao)
    assert_same_solution(collect(h), collect(co))


@pytest.mark.parametrize("n,m", [(12, 3), (64, 16)])
def test_wide_model_with_constant_jacobians(n, m):
    """A linear time-invariant plant on the wide-model path: the code generator flags the Jacobians as constants
    (ILQR_JAC_CONST) and the engine keeps ONE staged Jacobian block for the whole batch, written when the workspace is
    created -- gradients! has no dynamics part left.  n = 12: register-tiled DFMA Riccati loops, thread-per-problem
    forward kernel; n = 64: DMMA tiles and k_forward_wp.  Bit for bit against the oracle, MPC step included."""
    B, T = 5, 10
    model = problems.lq_invariant(n, m)
    assert "#define ILQR_JAC_CONST 1" in (model.header() if callable(model.header) else model.header)
    rng = np.random.default_rng(17)
    w = np.sin(0.3 * np.arange(T)[None, :, None] + rng.uniform(0, 6, (B, 1, n)))
    x1 = rng.standard_normal((B, n))
    ubar = 0.3 * rng.standard_normal((B, T - 1, m))
    kw = dict(objective_tolerance=0.0, lagrangian_gradient_tolerance=0.0, max_iterations=3)
    co = COracle(model, T, B, options=COptions.default(**kw))
    go = capi.default_options()
    for k, v in kw.items():
        setattr(go, k, v)
    h = capi.Handle(build.model_library(model), T, model.n, model.m, model.p, model.cs, model.ct, B, options=go)
    co.set_parameters(w); h.set_parameters(w)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))
    assert collect(h)["stats"]["iterations"].min() >= 2 and collect(h)["stats"]["flags"].max() == 0
    ao, xo = co.mpc_step(); ag, xg = h.mpc_step()
    np.testing.assert_array_equal(ag, ao)
    assert_same_solution(collect(h), collect(co))


def test_explicit_derivative_model_solves_like_the_traced_one():
    """A model whose dynamics and terminal constraint are given as C snippets (explicit-derivative constructors,
    src/dynamics.jl:55-60, src/constraints.jl:54-64) through the engine: same bits as the oracle compiled from the same
    header, and the same solution as the traced particle model."""
    from ilqr_b200 import Cost, constraint_from_c, dot, dynamics_from_c
    from ilqr_b200.api import Model
    dyn = dynamics_from_c("y[0] = x[0] + x[1];\ny[1] = x[1] + u[0];", "fx[0] = 1.0; fx[1] = 0.0; fx[2] = 1.0; fx[3] = 1.0;",
                          "fu[0] = 0.0; fu[1] = 1.0;", 2, 1)
    goal = constraint_from_c("c[0] = x[0] - 1.0; c[1] = x[1];", "cx[0] = 1.0; cx[1] = 0.0; cx[2] = 0.0; cx[3] = 1.0;", "", 2, 2, 0)
    m = Model("particle_c", dyn, Cost(lambda x, u: 0.1 * dot(x, x) + 0.1 * dot(u, u), 2, 1), Cost(lambda x, u: 0.1 * dot(x, x), 2, 0), None, goal)
    B, T = 20, 11
    _, x1, ubar = inputs("particle", B, T, seed=17)
    co = COracle(m, T, B)
    h = capi.Handle(build.model_library(m), T, m.n, m.m, m.p, m.cs, m.ct, B)
    solve_both(co, h, x1, ubar)
    assert_same_solution(collect(h), collect(co))
    _, _, _, co2, h2 = make_pair("particle", B, T, seed=17)
    solve_both(co2, h2, x1, ubar)
    np.testing.assert_array_equal(collect(h)["x"], collect(h2)["x"])
