"""Problem-definition layer: Dynamics / Cost / Constraint.

Host-side mirror of the reference's L1 layer (same names, argument meaning and
defaults): /root/reference/src/dynamics.jl:1-34, src/costs.jl:1-44,
src/constraints.jl:1-52.  Differences, all forced by the host language:

* user functions are traced with sympy instead of Symbolics.jl (codegen.py);
* ``indices_inequality`` is 0-based (Python), the reference is 1-based (Julia);
* the ``evaluate`` / ``jacobian_*`` / ``gradient_*`` / ``hessian_*`` members are
  in-place numpy callables ``fn(out, x, u, w)`` for inspecting a model on the host
  (what test/dynamics.jl, test/objective.jl, test/constraints.jl exercise).  The
  solve path never calls them: it runs the emitted C inside the CUDA kernels.
"""
from __future__ import annotations

import numpy as np
import sympy as sp

from . import codegen as cg


class Dynamics:
    """x+ = f(x, u[, w]).  Mirrors ``Dynamics(f, num_state, num_action; num_parameter)``
    (src/dynamics.jl:16-34)."""

    def __init__(self, f, num_state: int, num_action: int, num_parameter: int = 0, _traced=None):
        y, x, u, w = _traced if _traced is not None else cg.trace(f, num_state, num_action, num_parameter)
        self.x, self.u, self.w = x, u, w
        self.y = y
        self.fx = cg.jacobian(y, x)
        self.fu = cg.jacobian(y, u)
        self.num_next_state = len(y)
        self.num_state, self.num_action, self.num_parameter = num_state, num_action, num_parameter
        ny = self.num_next_state
        self.evaluate = cg.lambdify_inplace(y, (ny,), x, u, w)
        self.jacobian_state = cg.lambdify_inplace(cg._colmajor(self.fx), (ny, num_state), x, u, w)
        self.jacobian_action = cg.lambdify_inplace(cg._colmajor(self.fu), (ny, num_action), x, u, w)
        # scratch caches, as in src/dynamics.jl:31-33
        self.evaluate_cache = np.zeros(ny)
        self.jacobian_state_cache = np.zeros((ny, num_state))
        self.jacobian_action_cache = np.zeros((ny, num_action))


class _RawC:
    """What the explicit-derivative constructors store instead of traced expressions: C statement bodies (strings) that
    write the flat column-major outputs from ``x``, ``u``, ``w`` -- the engine's counterpart of the reference's
    user-provided in-place callables ``fn(out, x, u, w)`` (src/dynamics.jl:55-60, src/constraints.jl:54-64)."""

    def __init__(self, **bodies):
        self.__dict__.update(bodies)


def _no_host_callable(*_a, **_k):
    raise NotImplementedError("a model given as C snippets has no host-side callable; evaluate it through the compiled "
                              "model (oracle.c_oracle.CModelFns in the tests)")


def dynamics_from_c(evaluate: str, jacobian_state: str, jacobian_action: str, num_state: int, num_action: int,
                    num_parameter: int = 0) -> "Dynamics":
    """Dynamics(f, fx, fu, num_next_state, num_state, num_action, num_parameter) -- src/dynamics.jl:55-60 -- with the three
    functions given as C statement bodies writing ``y[i]``, ``fx[i + j*n]`` (column-major n x n) and ``fu[i + j*n]``
    (n x m) from ``x[]``, ``u[]``, ``w[]``.  They are compiled as they are into the kernels (and the oracle): nothing is
    traced or differentiated.  Use ``ilqr_fma`` / ``ilqr_sin`` / ``ilqr_cos`` for host/device bit-reproducibility."""
    d = Dynamics.__new__(Dynamics)
    d.x = d.u = d.w = d.y = d.fx = d.fu = None
    d.raw_c = _RawC(evaluate=evaluate, jacobian_state=jacobian_state, jacobian_action=jacobian_action)
    d.num_next_state = d.num_state = int(num_state)
    d.num_action, d.num_parameter = int(num_action), int(num_parameter)
    d.evaluate = d.jacobian_state = d.jacobian_action = _no_host_callable
    d.evaluate_cache = np.zeros(num_state)
    d.jacobian_state_cache = np.zeros((num_state, num_state))
    d.jacobian_action_cache = np.zeros((num_state, num_action))
    return d


def constraint_from_c(evaluate: str, jacobian_state: str, jacobian_action: str, num_constraint: int, num_state: int,
                      num_action: int, indices_inequality=(), num_parameter: int = 0) -> "Constraint":
    """Constraint(f, fx, fu, num_constraint, num_state, num_action; indices_inequality, num_parameter) --
    src/constraints.jl:54-64 -- as C statement bodies writing ``c[i]``, ``cx[i + j*nc]`` and ``cu[i + j*nc]``."""
    k = Constraint.__new__(Constraint)
    k.x = k.u = k.w = k.c = k.cx = k.cu = None
    k.raw_c = _RawC(evaluate=evaluate, jacobian_state=jacobian_state, jacobian_action=jacobian_action)
    k.num_constraint, k.num_state, k.num_action, k.num_parameter = int(num_constraint), int(num_state), int(num_action), int(num_parameter)
    k.indices_inequality = sorted(int(i) for i in indices_inequality)
    k.evaluate = k.jacobian_state = k.jacobian_action = _no_host_callable
    k.evaluate_cache = np.zeros(num_constraint)
    k.jacobian_state_cache = np.zeros((num_constraint, num_state))
    k.jacobian_action_cache = np.zeros((num_constraint, num_action))
    return k


class Cost:
    """Scalar stage / terminal cost.  Mirrors ``Cost(f, num_state, num_action; num_parameter)``
    (src/costs.jl:17-44); a terminal cost is built with ``num_action = 0``
    (examples/acrobot.jl:103)."""

    def __init__(self, f, num_state: int, num_action: int, num_parameter: int = 0, _traced=None):
        y, x, u, w = _traced if _traced is not None else cg.trace(f, num_state, num_action, num_parameter)
        if len(y) != 1:
            raise ValueError("Cost function must return a scalar")
        self.x, self.u, self.w = x, u, w
        self.g = y[0]
        self.gx = [sp.diff(self.g, v) for v in x]
        self.gu = [sp.diff(self.g, v) for v in u]
        self.gxx = cg.jacobian(self.gx, x)
        self.guu = cg.jacobian(self.gu, u)
        self.gux = cg.jacobian(self.gu, x)
        self.num_state, self.num_action, self.num_parameter = num_state, num_action, num_parameter
        n, m = num_state, num_action
        self.evaluate = cg.lambdify_inplace([self.g], (1,), x, u, w)
        self.gradient_state = cg.lambdify_inplace(self.gx, (n,), x, u, w)
        self.gradient_action = cg.lambdify_inplace(self.gu, (m,), x, u, w)
        self.hessian_state_state = cg.lambdify_inplace(cg._colmajor(self.gxx), (n, n), x, u, w)
        self.hessian_action_action = cg.lambdify_inplace(cg._colmajor(self.guu), (m, m), x, u, w)
        self.hessian_action_state = cg.lambdify_inplace(cg._colmajor(self.gux), (m, n), x, u, w)
        self.evaluate_cache = np.zeros(1)
        self.gradient_state_cache = np.zeros(n)
        self.gradient_action_cache = np.zeros(m)
        self.hessian_state_state_cache = np.zeros((n, n))
        self.hessian_action_action_cache = np.zeros((m, m))
        self.hessian_action_state_cache = np.zeros((m, n))


class Constraint:
    """c(x, u[, w]) (= 0, or <= 0 for rows in ``indices_inequality``).  Mirrors
    ``Constraint(f, num_state, num_action; indices_inequality, num_parameter)`` and the
    empty ``Constraint()`` (src/constraints.jl:17-52)."""

    def __init__(self, f=None, num_state: int = 0, num_action: int = 0,
                 indices_inequality=(), num_parameter: int = 0, _traced=None):
        if f is None and _traced is None:  # Constraint()  -- src/constraints.jl:45-52
            self.x, self.u, self.w = [], [], []
            self.c = []
            self.cx = sp.zeros(0, 0)
            self.cu = sp.zeros(0, 0)
            self.num_constraint = 0
            self.num_state = self.num_action = self.num_parameter = 0
            self.indices_inequality = []
            self.evaluate = self.jacobian_state = self.jacobian_action = lambda out, x, u=(), w=None: None
            self.evaluate_cache = np.zeros(0)
            self.jacobian_state_cache = np.zeros((0, 0))
            self.jacobian_action_cache = np.zeros((0, 0))
            return
        y, x, u, w = _traced if _traced is not None else cg.trace(f, num_state, num_action, num_parameter)
        self.x, self.u, self.w = x, u, w
        self.c = y
        self.cx = cg.jacobian(y, x)
        self.cu = cg.jacobian(y, u)
        self.num_constraint = len(y)
        self.num_state, self.num_action, self.num_parameter = num_state, num_action, num_parameter
        self.indices_inequality = sorted(int(i) for i in indices_inequality)
        for i in self.indices_inequality:
            if not 0 <= i < self.num_constraint:
                raise ValueError(f"indices_inequality entry {i} out of range (0-based)")
        nc = self.num_constraint
        self.evaluate = cg.lambdify_inplace(y, (nc,), x, u, w)
        self.jacobian_state = cg.lambdify_inplace(cg._colmajor(self.cx), (nc, num_state), x, u, w)
        self.jacobian_action = cg.lambdify_inplace(cg._colmajor(self.cu), (nc, num_action), x, u, w)
        self.evaluate_cache = np.zeros(nc)
        self.jacobian_state_cache = np.zeros((nc, num_state))
        self.jacobian_action_cache = np.zeros((nc, num_action))


class Model:
    """One compiled problem family: a stage (t < T) and a terminal (t = T) kind of
    each function.  The reference takes per-t vectors
    (src/solver.jl:28-30); its examples always repeat one object for t < T
    (examples/acrobot.jl:91-92), which is the case this engine compiles."""

    def __init__(self, name: str, dynamics: Dynamics, cost_stage: Cost, cost_terminal: Cost,
                 con_stage: Constraint | None = None, con_terminal: Constraint | None = None):
        self.name = name
        self.dynamics = dynamics
        self.cost_stage, self.cost_terminal = cost_stage, cost_terminal
        self.con_stage = con_stage if con_stage is not None else Constraint()
        self.con_terminal = con_terminal if con_terminal is not None else Constraint()
        d = dynamics
        if d.num_next_state != d.num_state:
            raise NotImplementedError("one Dynamics with num_next_state != num_state: give the per-step lists to Solver(...), "
                                      "which embeds time-varying dimensions in the largest ones (merge_stage_variants)")
        for obj, what in ((cost_stage, "stage cost"), (self.con_stage, "stage constraint")):
            if getattr(obj, "num_state", 0) and (obj.num_state != d.num_state or obj.num_action != d.num_action):
                raise ValueError(f"{what}: dimensions do not match the dynamics")
        for obj, what in ((cost_terminal, "terminal cost"), (self.con_terminal, "terminal constraint")):
            if getattr(obj, "num_state", 0) and (obj.num_state != d.num_state or obj.num_action != 0):
                raise ValueError(f"{what}: must have num_state = n and num_action = 0")
        for obj in (cost_stage, cost_terminal, self.con_stage, self.con_terminal):
            if getattr(obj, "num_parameter", 0) not in (0, d.num_parameter):
                raise ValueError("all functions of a model must share num_parameter (or take none)")
        self.n, self.m, self.p = d.num_state, d.num_action, d.num_parameter
        self.cs, self.ct = self.con_stage.num_constraint, self.con_terminal.num_constraint
        self._header = None

    @property
    def header(self) -> str:
        if self._header is None:
            self._header = cg.emit_header(self.name, self.dynamics, self.cost_stage, self.cost_terminal,
                                          self.con_stage, self.con_terminal)
        return self._header

    @property
    def hash(self) -> str:
        return cg.header_hash(self.header)

    @property
    def constrained(self) -> bool:
        return self.cs + self.ct > 0


# ----------------------------------------------------------------------------- time-varying stage functions
def _select(kind_symbol, variants):
    """kind == 0 ? variants[0] : kind == 1 ? variants[1] : ... (entry-wise); identical entries stay as they are"""
    out = []
    for entries in zip(*variants):
        first = entries[0]
        if all(e == first for e in entries[1:]):
            out.append(first)
        else:
            pieces = [(e, sp.Eq(kind_symbol, k)) for k, e in enumerate(entries[:-1])] + [(entries[-1], True)]
            out.append(sp.Piecewise(*pieces))
    return out


def merge_stage_variants(dynamics, costs, constraints):
    """The reference takes one Dynamics / Cost / Constraint object PER TIME STEP (src/solver.jl:28-30, README.md:26
    "time-varying").  The engine compiles one stage function of each kind, so distinct per-step objects of equal
    are merged into ONE function that selects its variant by an extra trailing parameter w[p] (the variant index of
    the step, filled in by the Solver): same values, one compiled model.

    Time-varying DIMENSIONS (num_state / num_action / num_next_state differing between steps, src/dynamics.jl:5) are
    embedded in the largest ones: a shorter next state is continued with zeros, a shorter action vector with components
    that no dynamics reads and that cost u_a^2 / 2 (which keeps Quu positive definite with an identity block behind the
    step's own; their gains and feed-forward terms are exact zeros).  Every sum of the solve then only gains exact zeros
    AFTER the step's own terms, so the padded solve reproduces the unpadded one (tests/test_gpu_parity.py::
    test_time_varying_dimensions, against the literal oracle working on the true per-step shapes).

    Returns (dynamics, cost, constraint, kinds per step, dims) with dims = None or (states per step 1..T, actions per
    step 1..T-1)."""
    steps = len(dynamics)
    triples, kinds = [], []
    for t in range(steps):
        tr = (dynamics[t], costs[t], constraints[t])
        for k, have in enumerate(triples):
            if all(a is b for a, b in zip(tr, have)):
                kinds.append(k)
                break
        else:
            kinds.append(len(triples))
            triples.append(tr)
    d0, c0, k0 = triples[0]
    p = d0.num_parameter
    n = max(max(d.num_state, d.num_next_state) for d, _, _ in triples)
    m = max(d.num_action for d, _, _ in triples)
    ns = [d.num_state for d in dynamics] + [dynamics[-1].num_next_state]
    ms = [d.num_action for d in dynamics]
    ragged = any(v != n for v in ns) or any(v != m for v in ms)
    for t in range(steps - 1):
        if dynamics[t].num_next_state != dynamics[t + 1].num_state:
            raise AssertionError(f"dynamics[{t}].num_next_state != dynamics[{t + 1}].num_state")
    for d, c, k in triples:
        if (c.num_state, c.num_action) != (d.num_state, d.num_action):
            raise AssertionError("a step's cost must have its dynamics' num_state and num_action")
        if getattr(k, "num_constraint", 0) and (k.num_state, k.num_action) != (d.num_state, d.num_action):
            raise AssertionError("a step's constraint must have its dynamics' num_state and num_action")
        if getattr(d, "raw_c", None) is not None or getattr(k, "raw_c", None) is not None:
            raise NotImplementedError("per-step objects built from C snippets cannot be merged")
        if d.num_parameter != p or c.num_parameter not in (0, p) or getattr(k, "num_parameter", 0) not in (0, p):
            raise NotImplementedError("stage functions with different num_parameter")
        if k.num_constraint != k0.num_constraint or list(k.indices_inequality) != list(k0.indices_inequality):
            raise NotImplementedError("stage constraints whose row count or inequality rows differ between steps")
    x, u = cg.SymVec(cg._symbols("x", n)), cg.SymVec(cg._symbols("u", m))
    w = cg.SymVec(cg._symbols("w", p + 1))
    sel = w[p]
    half = sp.Rational(1, 2)
    dyn = Dynamics(None, n, m, p + 1, _traced=(_select(sel, [list(d.y) + [sp.Integer(0)] * (n - len(d.y)) for d, _, _ in triples]), x, u, w))
    cost = Cost(None, n, m, p + 1,
                _traced=(_select(sel, [[c.g + sum((half * u[a] ** 2 for a in range(c.num_action, m)), sp.Integer(0))] for _, c, _ in triples]), x, u, w))
    if k0.num_constraint:
        con = Constraint(None, n, m, k0.indices_inequality, p + 1, _traced=(_select(sel, [list(k.c) for _, _, k in triples]), x, u, w))
    else:
        con = Constraint()
    return dyn, cost, con, kinds, ((ns, ms) if ragged else None)


def with_extra_parameter(obj, kind: str, p_new: int, num_state: int | None = None):
    """re-trace a terminal Cost / Constraint so that it shares the merged model's parameter count (the extra entry is
    unused) and, with time-varying dimensions, its padded state"""
    n = num_state if num_state is not None else obj.num_state
    x, u, w = cg.SymVec(cg._symbols("x", n)), cg.SymVec(cg._symbols("u", 0)), cg.SymVec(cg._symbols("w", p_new))
    if kind == "cost":
        return Cost(None, n, 0, p_new, _traced=([obj.g], x, u, w))
    if obj.num_constraint == 0:
        return obj
    return Constraint(None, n, 0, obj.indices_inequality, p_new, _traced=(list(obj.c), x, u, w))
