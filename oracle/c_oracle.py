"""ctypes binding of the C oracle (oracle/ilqr_oracle.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from build_oracle import build_for_model  # noqa: E402


class COptions(C.Structure):
    """struct ilqr_options of include/ilqr_cuda.h"""
    _fields_ = [("line_search", C.c_int32), ("max_iterations", C.c_int32), ("max_dual_updates", C.c_int32),
                ("reset_cache", C.c_int32), ("verbose", C.c_int32), ("reserved", C.c_int32),
                ("min_step_size", C.c_double), ("objective_tolerance", C.c_double),
                ("lagrangian_gradient_tolerance", C.c_double), ("constraint_tolerance", C.c_double),
                ("constraint_norm", C.c_double), ("initial_constraint_penalty", C.c_double),
                ("scaling_penalty", C.c_double), ("max_penalty", C.c_double)]

    @classmethod
    def default(cls, **kw):
        o = cls(0, 100, 10, 0, 0, 0, 1e-5, 1e-3, 1e-3, 5e-3, float("inf"), 1.0, 10.0, 1e8)
        for k, v in kw.items():
            if k == "line_search" and isinstance(v, str):
                v = {"armijo": 0, "none": 1}[v]
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, v)
        return o


def _p(a, typ=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(typ))


class COracle:
    """Batched front of the C oracle with the same host-buffer layouts as the C ABI
    ([problem][time][component])."""

    def __init__(self, model, T: int, batch: int, options: COptions | None = None, history_cap: int = 1000):
        self.lib = C.CDLL(build_for_model(model))
        L = self.lib
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(COptions)]
        L.oracle_model_hash.restype = C.c_char_p
        for f in ("oracle_destroy", "oracle_set_options", "oracle_initialize_controls", "oracle_initialize_states",
                  "oracle_set_parameters", "oracle_rollout", "oracle_solve", "oracle_mpc_step", "oracle_get_trajectory",
                  "oracle_get_stats", "oracle_get_history", "oracle_get_duals", "oracle_get_policy"):
            getattr(L, f).restype = None
        assert L.oracle_model_hash().decode() == model.hash
        self.model, self.T, self.B, self.cap = model, T, batch, history_cap
        self.n, self.m, self.p, self.cs, self.ct = model.n, model.m, model.p, model.cs, model.ct
        self.options = options or COptions.default()
        self.h = C.c_void_p(L.oracle_create(T, batch, history_cap, C.byref(self.options)))

    def __del__(self):
        try:
            self.lib.oracle_destroy(self.h)
        except Exception:
            pass

    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def set_options(self, options):
        self.options = options
        self.lib.oracle_set_options(self.h, C.byref(options))

    def _chk(self, a, shape):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == shape, (a.shape, shape)
        return a

    def initialize_controls(self, u):
        u = self._chk(u, (self.B, self.T - 1, self.m))
        self.lib.oracle_initialize_controls(self.h, _p(u))

    def initialize_states(self, x):
        x = self._chk(x, (self.B, self.T, self.n))
        self.lib.oracle_initialize_states(self.h, _p(x))

    def set_parameters(self, w):
        w = self._chk(w, (self.B, self.T, self.p))
        self.lib.oracle_set_parameters(self.h, _p(w))

    def rollout(self, x1, u):
        x1 = self._chk(x1, (self.B, self.n))
        u = self._chk(u, (self.B, self.T - 1, self.m))
        out = np.zeros((self.B, self.T, self.n))
        self.lib.oracle_rollout(self.h, _p(x1), _p(u), _p(out))
        return out

    def solve(self, nthreads: int = 0):
        self.lib.oracle_solve(self.h, C.c_int(nthreads))

    def mpc_step(self, nthreads: int = 0):
        au = np.zeros((self.B, self.m))
        xn = np.zeros((self.B, self.n))
        self.lib.oracle_mpc_step(self.h, C.c_int(nthreads), _p(au), _p(xn))
        return au, xn

    def get_trajectory(self, current: bool = False):
        x = np.zeros((self.B, self.T, self.n))
        u = np.zeros((self.B, self.T - 1, self.m))
        self.lib.oracle_get_trajectory(self.h, _p(x), _p(u), C.c_int(int(current)))
        return x, u

    def get_stats(self):
        B = self.B
        it = np.zeros(B, np.int32); st = np.zeros(B, np.uint8); J = np.zeros(B); mv = np.zeros(B); ss = np.zeros(B)
        fl = np.zeros(B, np.uint32)
        self.lib.oracle_get_stats(self.h, _p(it, C.c_int32), _p(st, C.c_uint8), _p(J), _p(mv), _p(ss), _p(fl, C.c_uint32))
        return dict(iterations=it, status=st, objective=J, max_violation=mv, step_size=ss, flags=fl)

    def get_history(self, cap: int | None = None):
        cap = cap or self.cap
        B = self.B
        cost = np.zeros((B, cap)); gn = np.zeros((B, cap)); mv = np.zeros((B, cap)); ss = np.zeros((B, cap))
        outer = np.zeros((B, cap), np.int32); st = np.zeros((B, cap), np.uint8)
        self.lib.oracle_get_history(self.h, C.c_int(cap), _p(cost), _p(gn), _p(mv), _p(ss), _p(outer, C.c_int32), _p(st, C.c_uint8))
        return dict(cost=cost, gradient_norm=gn, max_violation=mv, step_size=ss, outer=outer, status=st)

    def get_duals(self):
        rows = (self.T - 1) * self.cs + self.ct
        B = self.B
        lam = np.zeros((B, rows)); rho = np.zeros((B, rows)); c = np.zeros((B, rows)); a = np.zeros((B, rows), np.int32)
        self.lib.oracle_get_duals(self.h, _p(lam), _p(rho), _p(c), _p(a, C.c_int32))
        return dict(dual=lam, penalty=rho, violations=c, active_set=a)

    def get_policy(self):
        K = np.zeros((self.B, self.T - 1, self.m * self.n)); k = np.zeros((self.B, self.T - 1, self.m))
        self.lib.oracle_get_policy(self.h, _p(K), _p(k))
        return K, k


class CModelFns:
    """The generated C model functions, one call at a time, as reference-style in-place
    callables fn(out, x, u, w): lets oracle/ilqr_oracle.py run on exactly the model
    arithmetic the C oracle and the CUDA engine use."""

    def __init__(self, model):
        self.lib = C.CDLL(build_for_model(model))
        for f in ("oracle_eval_dyn", "oracle_eval_cost", "oracle_eval_con"):
            getattr(self.lib, f).restype = None
        self.model = model
        n, m, p, cs, ct = model.n, model.m, model.p, model.cs, model.ct
        self.n, self.m, self.p, self.cs, self.ct = n, m, p, cs, ct

    @staticmethod
    def _v(a, size):
        out = np.zeros(max(size, 1))
        if a is not None and size:
            flat = np.asarray(a, dtype=float).ravel()[:size]
            out[: flat.size] = flat  # a terminal stage passes an empty action
        return out

    def dyn(self, x, u, w=None):
        n, m, p = self.n, self.m, self.p
        y = np.zeros(n); fx = np.zeros(n * n); fu = np.zeros(max(n * m, 1))
        self.lib.oracle_eval_dyn(_p(self._v(x, n)), _p(self._v(u, m)), _p(self._v(w, p)), _p(y), _p(fx), _p(fu))
        return y, fx.reshape(n, n).T.copy(), fu[: n * m].reshape(m, n).T.copy()

    def cost(self, terminal, x, u=None, w=None):
        n, m, p = self.n, self.m, self.p
        g = np.zeros(1); gx = np.zeros(n); gu = np.zeros(max(m, 1)); gxx = np.zeros(n * n)
        guu = np.zeros(max(m * m, 1)); gux = np.zeros(max(m * n, 1))
        self.lib.oracle_eval_cost(C.c_int(int(terminal)), _p(self._v(x, n)), _p(self._v(u, m)), _p(self._v(w, p)),
                                  _p(g), _p(gx), _p(gu), _p(gxx), _p(guu), _p(gux))
        if terminal:
            return g[0], gx, None, gxx.reshape(n, n).T.copy(), None, None
        return (g[0], gx, gu[:m], gxx.reshape(n, n).T.copy(), guu[: m * m].reshape(m, m).T.copy(),
                gux[: m * n].reshape(n, m).T.copy())

    def con(self, terminal, x, u=None, w=None):
        n, m, p = self.n, self.m, self.p
        r = self.ct if terminal else self.cs
        c = np.zeros(max(r, 1)); cx = np.zeros(max(r * n, 1)); cu = np.zeros(max(r * m, 1))
        self.lib.oracle_eval_con(C.c_int(int(terminal)), _p(self._v(x, n)), _p(self._v(u, m)), _p(self._v(w, p)),
                                 _p(c), _p(cx), _p(cu))
        cxm = cx[: r * n].reshape(n, r).T.copy()
        cum = None if terminal else cu[: r * m].reshape(m, r).T.copy()
        return c[:r], cxm, cum

    # ---- reference-style object lists for OracleSolver
    def as_reference_objects(self, T):
        fns = self
        model = self.model

        class _Dyn:
            num_state, num_action, num_parameter, num_next_state = model.n, model.m, model.p, model.n
            evaluate_cache = np.zeros(model.n)
            jacobian_state_cache = np.zeros((model.n, model.n))
            jacobian_action_cache = np.zeros((model.n, model.m))

            @staticmethod
            def evaluate(out, x, u, w):
                out[...] = fns.dyn(x, u, w)[0]

            @staticmethod
            def jacobian_state(out, x, u, w):
                out[...] = fns.dyn(x, u, w)[1]

            @staticmethod
            def jacobian_action(out, x, u, w):
                out[...] = fns.dyn(x, u, w)[2]

        def mk_cost(terminal):
            n, m = model.n, (0 if terminal else model.m)

            class _Cost:
                evaluate_cache = np.zeros(1)
                gradient_state_cache = np.zeros(n)
                gradient_action_cache = np.zeros(m)
                hessian_state_state_cache = np.zeros((n, n))
                hessian_action_action_cache = np.zeros((m, m))
                hessian_action_state_cache = np.zeros((m, n))

                @staticmethod
                def evaluate(out, x, u, w):
                    out[0] = fns.cost(terminal, x, u, w)[0]

                @staticmethod
                def gradient_state(out, x, u, w):
                    out[...] = fns.cost(terminal, x, u, w)[1]

                @staticmethod
                def gradient_action(out, x, u, w):
                    out[...] = fns.cost(terminal, x, u, w)[2]

                @staticmethod
                def hessian_state_state(out, x, u, w):
                    out[...] = fns.cost(terminal, x, u, w)[3]

                @staticmethod
                def hessian_action_action(out, x, u, w):
                    out[...] = fns.cost(terminal, x, u, w)[4]

                @staticmethod
                def hessian_action_state(out, x, u, w):
                    out[...] = fns.cost(terminal, x, u, w)[5]

            return _Cost

        def mk_con(terminal):
            r = model.ct if terminal else model.cs
            src = model.con_terminal if terminal else model.con_stage
            n, m = model.n, (0 if terminal else model.m)

            class _Con:
                num_constraint = r
                indices_inequality = list(src.indices_inequality)
                evaluate_cache = np.zeros(r)
                jacobian_state_cache = np.zeros((r, n))
                jacobian_action_cache = np.zeros((r, m))

                @staticmethod
                def evaluate(out, x, u, w):
                    out[...] = fns.con(terminal, x, u, w)[0]

                @staticmethod
                def jacobian_state(out, x, u, w):
                    out[...] = fns.con(terminal, x, u, w)[1]

                @staticmethod
                def jacobian_action(out, x, u, w):
                    out[...] = fns.con(terminal, x, u, w)[2]

            return _Con

        dyn = [_Dyn] * (T - 1)
        obj = [mk_cost(False)] * (T - 1) + [mk_cost(True)]
        con = ([mk_con(False)] * (T - 1) + [mk_con(True)]) if model.constrained else None
        return dyn, obj, con
