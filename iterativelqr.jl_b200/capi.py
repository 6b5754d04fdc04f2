"""ctypes binding of libilqr_cuda.so (include/ilqr_cuda.h) -- what a Julia shim does with ccall.

The library is loaded from the in-tree build directory; a missing library or model plug-in
is an error (there is no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build

ABI_VERSION = 2


class IlqrOptions(C.Structure):
    _fields_ = [("line_search", C.c_int32), ("max_iterations", C.c_int32), ("max_dual_updates", C.c_int32),
                ("reset_cache", C.c_int32), ("verbose", C.c_int32), ("reserved", C.c_int32),
                ("min_step_size", C.c_double), ("objective_tolerance", C.c_double),
                ("lagrangian_gradient_tolerance", C.c_double), ("constraint_tolerance", C.c_double),
                ("constraint_norm", C.c_double), ("initial_constraint_penalty", C.c_double),
                ("scaling_penalty", C.c_double), ("max_penalty", C.c_double)]


class IlqrDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("T", C.c_int32), ("n", C.c_int32), ("m", C.c_int32), ("p", C.c_int32),
                ("c_s", C.c_int32), ("c_T", C.c_int32), ("batch", C.c_int32), ("device", C.c_int32),
                ("history_cap", C.c_int32), ("model_library", C.c_char_p)]


class IlqrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libilqr_cuda error {code}: {msg}")
        self.code = code


_PD = C.POINTER(C.c_double)
_PI32 = C.POINTER(C.c_int32)
_PU8 = C.POINTER(C.c_uint8)
_PU32 = C.POINTER(C.c_uint32)
_PI64 = C.POINTER(C.c_int64)

_SIGNATURES = {
    "ilqr_options_default": (None, [C.POINTER(IlqrOptions)]),
    "ilqr_create": (C.c_int, [C.POINTER(IlqrDesc), C.POINTER(IlqrOptions), C.POINTER(C.c_void_p)]),
    "ilqr_destroy": (None, [C.c_void_p]),
    "ilqr_last_error": (C.c_char_p, [C.c_void_p]),
    "ilqr_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ilqr_set_options": (C.c_int, [C.c_void_p, C.POINTER(IlqrOptions)]),
    "ilqr_initialize_controls": (C.c_int, [C.c_void_p, _PD]),
    "ilqr_initialize_states": (C.c_int, [C.c_void_p, _PD]),
    "ilqr_set_parameters": (C.c_int, [C.c_void_p, _PD]),
    "ilqr_initialize_controls_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ilqr_initialize_states_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ilqr_rollout": (C.c_int, [C.c_void_p, _PD, _PD, _PD]),
    "ilqr_solve": (C.c_int, [C.c_void_p]),
    "ilqr_solve_warm": (C.c_int, [C.c_void_p, _PD, _PD]),
    "ilqr_solve_outer": (C.c_int, [C.c_void_p, C.c_int32, _PI32]),
    "ilqr_solve_stream": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 11),
    "ilqr_solve_stream_host": (C.c_int, [C.c_void_p, C.c_int32, _PD, _PD, _PD, _PD, _PD, _PI32, _PU8, _PD, _PD, _PD, _PU32]),
    "ilqr_get_trajectory": (C.c_int, [C.c_void_p, _PD, _PD]),
    "ilqr_get_current_trajectory": (C.c_int, [C.c_void_p, _PD, _PD]),
    "ilqr_get_trajectory_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ilqr_get_stats": (C.c_int, [C.c_void_p, _PI32, _PU8, _PD, _PD, _PD, _PU32]),
    "ilqr_get_history": (C.c_int, [C.c_void_p, C.c_int32, _PD, _PD, _PD, _PD, _PI32, _PU8]),
    "ilqr_get_duals": (C.c_int, [C.c_void_p, _PD, _PD, _PD, _PI32]),
    "ilqr_get_policy": (C.c_int, [C.c_void_p, _PD, _PD]),
    "ilqr_mpc_step": (C.c_int, [C.c_void_p, _PD, _PD]),
    "ilqr_mpc_run": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ilqr_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "ilqr_get_counters": (C.c_int, [C.c_void_p, _PI64, _PI64, _PD, _PI64]),
    "ilqr_get_problem_ticks": (C.c_int, [C.c_void_p, _PI64]),
    "ilqr_get_compactions": (C.c_int, [C.c_void_p, _PI64]),
    "ilqr_comm_unique_id": (C.c_int, [C.c_char_p]),
    "ilqr_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p]),
    "ilqr_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ilqr_model_dims": (C.c_int, [C.c_char_p, _PI32, _PI32, _PI32, _PI32, _PI32]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def library_path() -> str:
    return build.front_library()


def lib():
    """Load libilqr_cuda.so (building it first if the in-tree copy is stale)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise IlqrError(-3, f"{path} missing; run __graft_entry__.build()")
        L = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # raises AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def default_options() -> IlqrOptions:
    o = IlqrOptions()
    lib().ilqr_options_default(C.byref(o))
    return o


class Handle:
    """Owning wrapper of an ``ilqr_handle*``."""

    def __init__(self, model_library: str, T, n, m, p, c_s, c_T, batch, device=0, history_cap=0, options=None):
        self._h = C.c_void_p()
        self.L = lib()
        self.T, self.n, self.m, self.p, self.c_s, self.c_T, self.B = T, n, m, p, c_s, c_T, batch
        self.history_cap = history_cap if history_cap > 0 else 1000
        desc = IlqrDesc(ABI_VERSION, T, n, m, p, c_s, c_T, batch, device, history_cap, model_library.encode())
        opt = options if options is not None else default_options()
        rc = self.L.ilqr_create(C.byref(desc), C.byref(opt), C.byref(self._h))
        if rc != 0:
            raise IlqrError(rc, (self.L.ilqr_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.L.ilqr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise IlqrError(rc, (self.L.ilqr_last_error(self._h) or b"").decode())

    @staticmethod
    def _arr(a, shape):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != tuple(shape):
            raise ValueError(f"expected array of shape {tuple(shape)}, got {a.shape}")
        return a

    @property
    def rows(self):
        return (self.T - 1) * self.c_s + self.c_T

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self.L.ilqr_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_options(self, opt: IlqrOptions):
        self._check(self.L.ilqr_set_options(self._h, C.byref(opt)))

    def initialize_controls(self, u):
        u = self._arr(u, (self.B, self.T - 1, self.m))
        self._check(self.L.ilqr_initialize_controls(self._h, _ptr(u, _PD)))

    def initialize_states(self, x):
        x = self._arr(x, (self.B, self.T, self.n))
        self._check(self.L.ilqr_initialize_states(self._h, _ptr(x, _PD)))

    def initialize_controls_device(self, d_u_ptr: int):
        self._check(self.L.ilqr_initialize_controls_device(self._h, C.c_void_p(d_u_ptr)))

    def initialize_states_device(self, d_x_ptr: int):
        self._check(self.L.ilqr_initialize_states_device(self._h, C.c_void_p(d_x_ptr)))

    def set_parameters(self, w):
        w = self._arr(w, (self.B, self.T, self.p))
        self._check(self.L.ilqr_set_parameters(self._h, _ptr(w, _PD)))

    def rollout(self, x1, u):
        x1 = self._arr(x1, (self.B, self.n))
        u = self._arr(u, (self.B, self.T - 1, self.m))
        out = np.empty((self.B, self.T, self.n))
        self._check(self.L.ilqr_rollout(self._h, _ptr(x1, _PD), _ptr(u, _PD), _ptr(out, _PD)))
        return out

    def solve(self):
        self._check(self.L.ilqr_solve(self._h))

    def solve_outer(self, restart: bool) -> int:
        """one outer (augmented-Lagrangian) iteration of every problem; returns how many wait for the next call"""
        n = C.c_int32()
        self._check(self.L.ilqr_solve_outer(self._h, int(bool(restart)), C.byref(n)))
        return n.value

    def solve_warm(self, x, u):
        x = self._arr(x, (self.B, self.T, self.n))
        u = self._arr(u, (self.B, self.T - 1, self.m))
        self._check(self.L.ilqr_solve_warm(self._h, _ptr(x, _PD), _ptr(u, _PD)))

    def solve_stream(self, n_problems: int, d_x: int, d_u: int, d_w: int = 0, out_x: int = 0, out_u: int = 0,
                     out_iterations: int = 0, out_status: int = 0, out_objective: int = 0, out_max_violation: int = 0,
                     out_step_size: int = 0, out_flags: int = 0):
        """All arguments are raw DEVICE pointers (ints); 0 = NULL."""
        ptrs = [C.c_void_p(p or None) for p in (d_x, d_u, d_w, out_x, out_u, out_iterations, out_status, out_objective,
                                                out_max_violation, out_step_size, out_flags)]
        self._check(self.L.ilqr_solve_stream(self._h, int(n_problems), *ptrs))

    def solve_stream_host(self, x, u, w=None, out_x=None, out_u=None):
        """n fresh problems (host arrays [n][T][n_state], [n][T-1][m]) streamed through the handle's slots."""
        n = x.shape[0]
        x = self._arr2(x, (n, self.T, self.n)); u = self._arr2(u, (n, self.T - 1, self.m))
        if self.p:
            w = self._arr2(w, (n, self.T, self.p))
        ox = out_x if out_x is not None else np.empty((n, self.T, self.n))
        ou = out_u if out_u is not None else np.empty((n, self.T - 1, self.m))
        it = np.zeros(n, np.int32); st = np.zeros(n, np.uint8); J = np.zeros(n); mv = np.zeros(n); ss = np.zeros(n)
        fl = np.zeros(n, np.uint32)
        self._check(self.L.ilqr_solve_stream_host(self._h, n, _ptr(x, _PD), _ptr(u, _PD), _ptr(w, _PD) if self.p else None,
                                                  _ptr(ox, _PD), _ptr(ou, _PD), _ptr(it, _PI32), _ptr(st, _PU8), _ptr(J, _PD),
                                                  _ptr(mv, _PD), _ptr(ss, _PD), _ptr(fl, _PU32)))
        return ox, ou, dict(iterations=it, status=st, objective=J, max_violation=mv, step_size=ss, flags=fl)

    @staticmethod
    def _arr2(a, shape):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != tuple(shape):
            raise ValueError(f"expected array of shape {tuple(shape)}, got {a.shape}")
        return a

    def get_trajectory(self, current=False, out_x=None, out_u=None):
        x = out_x if out_x is not None else np.empty((self.B, self.T, self.n))
        u = out_u if out_u is not None else np.empty((self.B, self.T - 1, self.m))
        fn = self.L.ilqr_get_current_trajectory if current else self.L.ilqr_get_trajectory
        self._check(fn(self._h, _ptr(x, _PD), _ptr(u, _PD)))
        return x, u

    def get_trajectory_device(self, d_x_ptr: int, d_u_ptr: int):
        self._check(self.L.ilqr_get_trajectory_device(self._h, C.c_void_p(d_x_ptr), C.c_void_p(d_u_ptr)))

    def get_stats(self):
        B = self.B
        it = np.zeros(B, np.int32); st = np.zeros(B, np.uint8); J = np.zeros(B); mv = np.zeros(B); ss = np.zeros(B)
        fl = np.zeros(B, np.uint32)
        self._check(self.L.ilqr_get_stats(self._h, _ptr(it, _PI32), _ptr(st, _PU8), _ptr(J, _PD), _ptr(mv, _PD),
                                          _ptr(ss, _PD), _ptr(fl, _PU32)))
        return dict(iterations=it, status=st, objective=J, max_violation=mv, step_size=ss, flags=fl)

    def get_history(self, cap=None):
        cap = cap or self.history_cap
        B = self.B
        cost = np.zeros((B, cap)); gn = np.zeros((B, cap)); mv = np.zeros((B, cap)); ss = np.zeros((B, cap))
        outer = np.zeros((B, cap), np.int32); st = np.zeros((B, cap), np.uint8)
        self._check(self.L.ilqr_get_history(self._h, cap, _ptr(cost, _PD), _ptr(gn, _PD), _ptr(mv, _PD), _ptr(ss, _PD),
                                            _ptr(outer, _PI32), _ptr(st, _PU8)))
        return dict(cost=cost, gradient_norm=gn, max_violation=mv, step_size=ss, outer=outer, status=st)

    def get_duals(self):
        B, rows = self.B, self.rows
        lam = np.zeros((B, rows)); rho = np.zeros((B, rows)); c = np.zeros((B, rows)); a = np.zeros((B, rows), np.int32)
        self._check(self.L.ilqr_get_duals(self._h, _ptr(lam, _PD), _ptr(rho, _PD), _ptr(c, _PD), _ptr(a, _PI32)))
        return dict(dual=lam, penalty=rho, violations=c, active_set=a)

    def get_policy(self):
        K = np.zeros((self.B, self.T - 1, self.m * self.n)); k = np.zeros((self.B, self.T - 1, self.m))
        self._check(self.L.ilqr_get_policy(self._h, _ptr(K, _PD), _ptr(k, _PD)))
        return K, k

    def mpc_step(self):
        au = np.zeros((self.B, self.m)); xn = np.zeros((self.B, self.n))
        self._check(self.L.ilqr_mpc_step(self._h, _ptr(au, _PD), _ptr(xn, _PD)))
        return au, xn

    def mpc_run(self, n_steps: int, d_applied_u: int = 0, d_x_next: int = 0, d_total_iterations: int = 0):
        """raw DEVICE pointers (ints); 0 = NULL"""
        self._check(self.L.ilqr_mpc_run(self._h, int(n_steps), C.c_void_p(d_applied_u or None), C.c_void_p(d_x_next or None),
                                        C.c_void_p(d_total_iterations or None)))

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        """join the gather communicator (ncclCommInitRank on this handle's device); ``unique_id`` from comm_unique_id()
        on rank 0, shipped to the other ranks by the host"""
        if len(unique_id) != 128:
            raise ValueError("unique_id must be 128 bytes")
        self._check(self.L.ilqr_comm_init(self._h, int(n_ranks), int(rank), unique_id))

    def gather(self, d_local: int, d_all: int, bytes_per_rank: int):
        """ncclAllGather on the handle's stream; raw DEVICE pointers (ints)"""
        self._check(self.L.ilqr_gather(self._h, C.c_void_p(d_local), C.c_void_p(d_all), int(bytes_per_rank)))

    def set_profiling(self, on: bool):
        self._check(self.L.ilqr_set_profiling(self._h, int(on)))

    def get_counters(self):
        ticks = C.c_int64(); launches = C.c_int64()
        ms = np.zeros(3); kl = np.zeros(3, np.int64)
        self._check(self.L.ilqr_get_counters(self._h, C.byref(ticks), C.byref(launches), _ptr(ms, _PD), _ptr(kl, _PI64)))
        pt = C.c_int64()
        self._check(self.L.ilqr_get_problem_ticks(self._h, C.byref(pt)))
        nc = C.c_int64()
        self._check(self.L.ilqr_get_compactions(self._h, C.byref(nc)))
        return dict(ticks=ticks.value, launches=launches.value, kernel_ms=ms, kernel_launches=kl, problem_ticks=pt.value,
                    compactions=nc.value)


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (rank 0 calls this; 128 bytes)"""
    buf = C.create_string_buffer(128)
    rc = lib().ilqr_comm_unique_id(buf)
    if rc != 0:
        raise IlqrError(rc, "ilqr_comm_unique_id failed (NCCL not loadable?)")
    return buf.raw


def model_dims(model_library: str):
    v = [C.c_int32() for _ in range(5)]
    rc = lib().ilqr_model_dims(model_library.encode(), *[C.byref(x) for x in v])
    if rc != 0:
        raise IlqrError(rc, (lib().ilqr_last_error(None) or b"").decode())
    return tuple(x.value for x in v)
