"""Solver layer: the reference's L2/L4 API names over the CUDA engine.

Mirrors /root/reference/src/solver.jl:4-66 (Solver, get_trajectory, current_trajectory,
initialize_controls!, initialize_states!), src/options.jl:1-14 (Options), src/rollout.jl:33-42
(rollout) and src/solve.jl:56-60, :131-143 (solve!).  Python has no ``!`` in identifiers, so
``initialize_controls!`` is ``initialize_controls`` etc.

One ``Solver`` holds a BATCH of independent problems of the same model (the reference
holds one); with ``batch=1`` the calls take and return the reference's shapes
(lists of per-time-step vectors), with ``batch=B`` arrays shaped [B][T][n] / [B][T-1][m].
Every numerical call goes through libilqr_cuda.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import build, capi
from .api import Constraint, Cost, Dynamics, Model


@dataclass
class Options:
    """src/options.jl:1-14, same names and defaults."""
    line_search: str = "armijo"
    max_iterations: int = 100
    max_dual_updates: int = 10
    min_step_size: float = 1.0e-5
    objective_tolerance: float = 1.0e-3
    lagrangian_gradient_tolerance: float = 1.0e-3
    constraint_tolerance: float = 5.0e-3
    constraint_norm: float = math.inf
    initial_constraint_penalty: float = 1.0
    scaling_penalty: float = 10.0
    max_penalty: float = 1.0e8
    reset_cache: bool = False
    verbose: bool = True

    def to_c(self) -> capi.IlqrOptions:
        if self.line_search not in ("armijo", "none"):
            raise ValueError("line_search must be 'armijo' or 'none'")
        return capi.IlqrOptions(0 if self.line_search == "armijo" else 1, int(self.max_iterations),
                                int(self.max_dual_updates), int(bool(self.reset_cache)), int(bool(self.verbose)), 0,
                                float(self.min_step_size), float(self.objective_tolerance),
                                float(self.lagrangian_gradient_tolerance), float(self.constraint_tolerance),
                                float(self.constraint_norm), float(self.initial_constraint_penalty),
                                float(self.scaling_penalty), float(self.max_penalty))


_zero_cost_cache: dict = {}


def _model_from_lists(dynamics, objective, constraints) -> tuple[Model, int]:
    """Collapse the reference's per-t vectors (src/solver.jl:28-30) into the stage/terminal
    pair the engine compiles.  The lists must repeat ONE object for t < T, as the reference's
    examples do (examples/acrobot.jl:92 "best to instantiate once")."""
    T = len(dynamics) + 1
    if len(objective) != T:
        raise AssertionError("length(dynamics) + 1 == length(costs)  (src/data/problem.jl:30)")
    if constraints is not None and len(constraints) != T:
        raise AssertionError("length(constraints) must be T")

    dyn_l, cost_l = list(dynamics), list(objective[:-1])
    cost_T = objective[-1]
    empty = Constraint()
    con_l = [empty] * (T - 1) if constraints is None else list(constraints[:-1])
    con_T = Constraint() if constraints is None else constraints[-1]
    uniform = all(o is dyn_l[0] for o in dyn_l) and all(o is cost_l[0] for o in cost_l) and all(o is con_l[0] for o in con_l) \
        and dyn_l[0].num_next_state == dyn_l[0].num_state
    key = tuple(id(o) for o in dyn_l + cost_l + con_l) + (id(cost_T), id(con_T)) if not uniform else \
        (id(dyn_l[0]), id(cost_l[0]), id(cost_T), id(con_l[0]) if constraints is not None else 0, id(con_T) if constraints is not None else 0)
    hit = _model_cache.get(key)
    if hit is not None:
        return hit[0], T, hit[1], hit[2]
    kinds = dims = None
    if uniform:
        dyn, cost_s, con_s = dyn_l[0], cost_l[0], (con_l[0] if constraints is not None else Constraint())
    else:  # distinct per-step objects (src/solver.jl:28-30): one merged stage function selecting by a trailing parameter
        from .api import merge_stage_variants, with_extra_parameter
        dyn, cost_s, con_s, kinds, dims = merge_stage_variants(dyn_l, cost_l, con_l)
        if cost_T.num_state != dyn_l[-1].num_next_state or (con_T.num_constraint and con_T.num_state != dyn_l[-1].num_next_state):
            raise AssertionError("terminal cost / constraint: num_state must be the last dynamics' num_next_state")
        cost_T = with_extra_parameter(cost_T, "cost", dyn.num_parameter, dyn.num_state)
        con_T = with_extra_parameter(con_T, "constraint", dyn.num_parameter, dyn.num_state)
    model = Model("user", dyn, cost_s, cost_T, con_s, con_T)
    model.name = f"user_{model.hash[:8]}"
    model._header = None  # name is part of the header text
    _model_cache[key] = (model, kinds, dims)
    _model_keepalive.append((dyn_l, cost_l, con_l, cost_T, con_T))
    return model, T, kinds, dims


_model_cache: dict = {}
_model_keepalive: list = []


class Solver:
    """``Solver(dynamics, objective[, constraints]; parameters, options)`` (src/solver.jl:11-46).

    Extra keyword arguments (engine-specific): ``batch`` (independent problems, default 1),
    ``device`` (CUDA ordinal), ``history_cap`` (iteration records kept per problem)."""

    def __init__(self, dynamics, objective=None, constraints=None, parameters=None, options=None,
                 batch: int = 1, device: int = 0, history_cap: int = 0, T: int | None = None):
        self.stage_kinds = None  # time-varying stage functions: variant index per step, carried in the last parameter
        self.dims = None         # time-varying dimensions: (states per step, actions per step); the engine works on the padded ones
        if isinstance(dynamics, Model):
            if T is None:
                raise ValueError("Solver(model, T=...) needs the horizon")
            model = dynamics
        else:
            model, T, self.stage_kinds, self.dims = _model_from_lists(dynamics, objective, constraints)
        self.model, self.T, self.batch = model, int(T), int(batch)
        self.options = options if options is not None else Options()
        self._lib_path = build.model_library(model)
        self.handle = capi.Handle(self._lib_path, self.T, model.n, model.m, model.p, model.cs, model.ct,
                                  self.batch, device=device, history_cap=history_cap, options=self.options.to_c())
        self._options_sent = self.options.to_c()
        if parameters is not None or self.stage_kinds is not None:
            self.set_parameters(parameters)

    # -- shape helpers: reference shapes for batch == 1, arrays otherwise
    def _in(self, a, steps, dim):
        if self.dims is not None:  # per-step vectors of the steps' own lengths ([n_t] or [batch][n_t]), zero-padded
            sizes = self.dims[0] if steps == self.T else self.dims[1]
            if len(a) != steps:
                raise ValueError(f"expected {steps} per-step vectors")
            out = np.zeros((self.batch, steps, dim))
            for t, v in enumerate(a):
                v = np.asarray(v, dtype=np.float64)
                if v.shape[-1] != sizes[t]:
                    raise ValueError(f"step {t}: expected {sizes[t]} components, got {v.shape[-1]}")
                out[:, t, :sizes[t]] = v
            return out
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 2 and self.batch == 1:
            a = a[None]
        if a.ndim == 2 and self.batch > 1 and a.shape == (steps, dim):
            a = np.broadcast_to(a, (self.batch, steps, dim))
        return np.ascontiguousarray(a.reshape(self.batch, steps, dim))

    def _out(self, a):
        if self.dims is not None:
            sizes = self.dims[0] if a.shape[1] == self.T else self.dims[1]
            return [(a[0, t, :sizes[t]] if self.batch == 1 else a[:, t, :sizes[t]]).copy() for t in range(a.shape[1])]
        return [a[0, t].copy() for t in range(a.shape[1])] if self.batch == 1 else a

    def _sync_options(self):
        c = self.options.to_c()
        if bytes(c) != bytes(self._options_sent):
            self.handle.set_options(c)
            self._options_sent = c

    def set_parameters(self, parameters):
        p = self.model.p
        if p == 0:
            return
        if self.stage_kinds is not None:  # the user's p - 1 parameters + the stage-variant selector
            pu = p - 1
            w = np.zeros((self.batch, self.T, p))
            if parameters is not None and pu > 0:
                if isinstance(parameters, (list, tuple)) and self.batch == 1:
                    for t, wt in enumerate(parameters):
                        wt = np.asarray(wt, dtype=float).ravel()
                        if wt.size:
                            w[0, t, :pu] = wt
                else:
                    wu = np.asarray(parameters, dtype=float)
                    if wu.ndim == 2:
                        wu = np.broadcast_to(wu, (self.batch,) + wu.shape)
                    w[:, :wu.shape[1], :pu] = wu
            w[:, :self.T - 1, pu] = np.asarray(self.stage_kinds, dtype=float)[None, :]
            self.handle.set_parameters(np.ascontiguousarray(w))
            return
        if isinstance(parameters, (list, tuple)) and len(parameters) in (self.T - 1, self.T) and self.batch == 1:
            w = np.zeros((1, self.T, p))
            for t, wt in enumerate(parameters):
                wt = np.asarray(wt, dtype=float).ravel()
                if wt.size:
                    w[0, t] = wt
        else:
            w = np.asarray(parameters, dtype=float)
            if w.ndim == 2:
                w = np.broadcast_to(w, (self.batch,) + w.shape)
            if w.shape[1] == self.T - 1:  # src/data/problem.jl:28: terminal entry defaults to empty
                w = np.concatenate([w, np.zeros((self.batch, 1, p))], axis=1)
        self.handle.set_parameters(np.ascontiguousarray(w))

    # -- diagnostics mirroring SolverData (src/data/solver.jl:4-18)
    @property
    def data(self):
        return self.handle.get_stats()

    def history(self, cap=None):
        return self.handle.get_history(cap)

    def close(self):
        self.handle.close()


def initialize_controls(solver: Solver, actions):
    """initialize_controls!(solver, actions) -- src/solver.jl:56-60"""
    solver.handle.initialize_controls(solver._in(actions, solver.T - 1, solver.model.m))


def initialize_states(solver: Solver, states):
    """initialize_states!(solver, states) -- src/solver.jl:62-66"""
    solver.handle.initialize_states(solver._in(states, solver.T, solver.model.n))


def get_trajectory(solver: Solver):
    """get_trajectory(solver) -- src/solver.jl:48-50 (nominal trajectory)"""
    x, u = solver.handle.get_trajectory()
    return solver._out(x), solver._out(u)


def current_trajectory(solver: Solver):
    """current_trajectory(solver) -- src/solver.jl:52-54"""
    x, u = solver.handle.get_trajectory(current=True)
    return solver._out(x), solver._out(u)


def _print_history(solver: Solver, start: int = 0):
    """The verbose printout of src/solve.jl:40-45, :106 (problem 0 of the batch), for the solve that has just finished:
    records are indexed by data.iterations, which an unconstrained solver never resets (src/solve.jl:137-139), so a repeated
    solve's records start at `start` = the counter before it."""
    st = solver.handle.get_stats()
    n = int(st["iterations"][0])
    h = solver.handle.get_history(min(max(n, 1), solver.handle.history_cap))
    last_outer = None
    it = 0
    for r in range(min(start, n), min(n, h["cost"].shape[1])):
        outer = int(h["outer"][0, r])
        if outer != last_outer:
            if outer > 0:
                print(f"  al iter: {outer}")
            last_outer, it = outer, 0
        it += 1
        print(f"iter:                  {it}\n"
              f"             cost:                  {h['cost'][0, r]}\n"
              f"             gradient_norm:         {h['gradient_norm'][0, r]}\n"
              f"             max_violation:         {h['max_violation'][0, r]}\n"
              f"             step_size:             {h['step_size'][0, r]}")


def solve(solver: Solver, states=None, actions=None, augmented_lagrangian_callback=None):
    """solve!(solver[, states, actions]; augmented_lagrangian_callback!) -- src/solve.jl:137-143 (+ warm start :56-60,
    :131-135).  The callback (src/solve.jl:88,125) is called with the solver after every dual update, the whole batch
    having made its update (problems that terminated earlier sit the call out); it may change parameters or options."""
    solver._sync_options()
    if (states is None) != (actions is None):
        raise TypeError("solve(solver, states, actions): give both or neither")
    printed_from = 0
    if solver.options.verbose and not solver.model.constrained and not solver.options.reset_cache:
        printed_from = int(solver.handle.get_stats()["iterations"][0])
    if augmented_lagrangian_callback is not None and solver.model.constrained:
        if states is not None:
            initialize_controls(solver, actions)
            initialize_states(solver, states)
        waiting = solver.handle.solve_outer(True)
        while waiting > 0:
            augmented_lagrangian_callback(solver)
            solver._sync_options()
            waiting = solver.handle.solve_outer(False)
    elif states is not None:
        solver.handle.solve_warm(solver._in(states, solver.T, solver.model.n),
                                 solver._in(actions, solver.T - 1, solver.model.m))
    else:
        solver.handle.solve()
    if solver.options.verbose:
        _print_history(solver, printed_from)
    return None


def solve_stream(solver: Solver, states, actions, parameters=None):
    """Continuous batching (ilqr_solve_stream_host): ``states`` [n][T][n_state] and ``actions`` [n][T-1][m] are n
    independent problems, each solved as a fresh ``Solver`` + ``solve!`` would (src/solver.jl:28-46,
    src/solve.jl:137-143), streamed through the solver's ``batch`` slots.  Returns (x, u, stats)."""
    solver._sync_options()
    x = np.ascontiguousarray(states, dtype=np.float64)
    u = np.ascontiguousarray(actions, dtype=np.float64)
    w = None if parameters is None else np.ascontiguousarray(parameters, dtype=np.float64)
    return solver.handle.solve_stream_host(x, u, w)


_rollout_solvers: dict = {}


def rollout(dynamics, initial_state, actions, parameters=None):
    """rollout(dynamics, initial_state, actions[, parameters]) -- src/rollout.jl:33-42.
    ``dynamics`` is the reference's per-t list (one repeated Dynamics).  Runs the open-loop
    rollout kernel; returns the list of T states (or [B][T][n] for batched input)."""
    dyn = dynamics[0]
    T = len(dynamics) + 1
    x1 = np.asarray(initial_state, dtype=float)
    batched = x1.ndim == 2
    B = x1.shape[0] if batched else 1
    if any(d is not dyn for d in dynamics):  # time-varying dynamics: through a Solver of the merged model (zero costs)
        key = tuple(id(d) for d in dynamics) + (B,)
        s = _rollout_solvers.get(key)
        if s is None:
            n, m, p = dyn.num_state, dyn.num_action, dyn.num_parameter
            zs = Cost((lambda x, u, w: 0 * x[0]) if p else (lambda x, u: 0 * x[0]), n, m, p)
            zT = Cost((lambda x, u, w: 0 * x[0]) if p else (lambda x, u: 0 * x[0]), n, 0, p)
            s = Solver(list(dynamics), [zs] * (T - 1) + [zT], batch=B, options=Options(verbose=False))
            _rollout_solvers[key] = s
        s.set_parameters(parameters)
        out = s.handle.rollout(x1.reshape(B, dyn.num_state), s._in(actions, T - 1, dyn.num_action))
        return out if batched else [out[0, t].copy() for t in range(T)]
    key = (id(dyn), T, B)
    s = _rollout_solvers.get(key)
    if s is None:
        zk = (dyn.num_state, dyn.num_action, dyn.num_parameter)
        zc = _zero_cost_cache.get(zk)
        if zc is None:
            n, m, p = zk
            zc = (Cost((lambda x, u, w: 0 * x[0]) if p else (lambda x, u: 0 * x[0]), n, m, p),
                  Cost((lambda x, u, w: 0 * x[0]) if p else (lambda x, u: 0 * x[0]), n, 0, p))
            _zero_cost_cache[zk] = zc
        s = Solver([dyn] * (T - 1), [zc[0]] * (T - 1) + [zc[1]], batch=B, options=Options(verbose=False))
        _rollout_solvers[key] = s
    if dyn.num_parameter:  # the cached solver must not keep an earlier call's parameters: the reference defaults to zeros
        s.set_parameters(parameters if parameters is not None else np.zeros((B, T, dyn.num_parameter)))
    out = s.handle.rollout(x1.reshape(B, dyn.num_state), s._in(actions, T - 1, dyn.num_action))
    return out if batched else [out[0, t].copy() for t in range(T)]
