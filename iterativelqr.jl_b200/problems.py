"""The reference's example / test problems, restated over the host API.

Constants follow /root/reference/examples/particle.jl:17-47,
examples/acrobot.jl:18-110 (= test/acrobot.jl:9-100), test/car.jl:10-61 and the
unit fixtures test/dynamics.jl:8-19.  ``lq_tracking`` is BASELINE.json config 4
(no reference analogue)."""
from __future__ import annotations

import functools
import math

import numpy as np
import sympy as sp

from .api import Constraint, Cost, Dynamics, Model
from .codegen import SymVec, dot, vcat


# ----------------------------------------------------------------------------- particle
def particle_discrete(x, u):
    # A = [1 1; 0 1], B = [0; 1]   (examples/particle.jl:17-21)
    return [1.0 * x[0] + 1.0 * x[1] + 0.0 * u[0], 0.0 * x[0] + 1.0 * x[1] + 1.0 * u[0]]


@functools.lru_cache(maxsize=None)
def particle() -> Model:
    n, m = 2, 1
    xT = [1.0, 0.0]
    return Model(
        "particle",
        Dynamics(particle_discrete, n, m),
        Cost(lambda x, u: 0.1 * dot(x, x) + 0.1 * dot(u, u), n, m),
        Cost(lambda x, u: 0.1 * dot(x, x), n, 0),
        Constraint(),
        Constraint(lambda x, u: x - xT, n, 0),
    )


# ----------------------------------------------------------------------------- acrobot
def acrobot_continuous(x, u):
    mass1, inertia1, length1, lengthcom1 = 1.0, 0.33, 1.0, 0.5
    mass2, inertia2, length2, lengthcom2 = 1.0, 0.33, 1.0, 0.5
    gravity, friction1, friction2 = 9.81, 0.1, 0.1
    sin, cos = sp.sin, sp.cos

    # M(q)  (examples/acrobot.jl:33-42)
    Ma = inertia1 + inertia2 + mass2 * length1 * length1 + 2.0 * mass2 * length1 * lengthcom2 * cos(x[1])
    Mb = inertia2 + mass2 * length1 * lengthcom2 * cos(x[1])
    Mc = inertia2
    # Minv via the 2x2 determinant (:44-51)
    a, b, c, d = Ma, Mb, Mb, Mc
    idet = 1.0 / (a * d - b * c)
    Minv = [[idet * d, idet * -b], [idet * -c, idet * a]]
    # tau(q)  (:53-61)
    ta = (-1.0 * mass1 * gravity * lengthcom1 * sin(x[0])
          - mass2 * gravity * (length1 * sin(x[0]) + lengthcom2 * sin(x[0] + x[1])))
    tb = -1.0 * mass2 * gravity * lengthcom2 * sin(x[0] + x[1])
    # C(x)  (:63-70)
    Ca = -2.0 * mass2 * length1 * lengthcom2 * sin(x[1]) * x[3]
    Cb = -1.0 * mass2 * length1 * lengthcom2 * sin(x[1]) * x[3]
    Cc = mass2 * length1 * lengthcom2 * sin(x[1]) * x[2]
    Cd = 0.0
    v = [x[2], x[3]]
    # qdd = Minv * (-C v + tau + B u - friction .* v), B = [0; 1]  (:79-80)
    r0 = -1.0 * (Ca * v[0] + Cb * v[1]) + ta + 0.0 * u[0] - friction1 * v[0]
    r1 = -1.0 * (Cc * v[0] + Cd * v[1]) + tb + 1.0 * u[0] - friction2 * v[1]
    qdd0 = Minv[0][0] * r0 + Minv[0][1] * r1
    qdd1 = Minv[1][0] * r0 + Minv[1][1] * r1
    return SymVec([x[2], x[3], qdd0, qdd1])


def acrobot_discrete(x, u):
    h = 0.1  # explicit midpoint (examples/acrobot.jl:85-88)
    return x + h * acrobot_continuous(x + 0.5 * h * acrobot_continuous(x, u), u)


@functools.lru_cache(maxsize=None)
def acrobot() -> Model:
    n, m = 4, 1
    xT = [math.pi, 0.0, 0.0, 0.0]
    return Model(
        "acrobot",
        Dynamics(acrobot_discrete, n, m),
        Cost(lambda x, u: 0.1 * dot(x[2:4], x[2:4]) + 0.1 * dot(u, u), n, m),
        Cost(lambda x, u: 0.1 * dot(x[2:4], x[2:4]), n, 0),
        Constraint(),
        Constraint(lambda x, u: x - xT, n, 0),
    )


# ----------------------------------------------------------------------------- car
def car_continuous(x, u):
    return SymVec([u[0] * sp.cos(x[2]), u[0] * sp.sin(x[2]), u[1]])


def car_discrete(x, u):
    h = 0.1  # test/car.jl:14-17
    return x + h * car_continuous(x + 0.5 * h * car_continuous(x, u), u)


@functools.lru_cache(maxsize=None)
def car() -> Model:
    n, m = 3, 2
    xT = [1.0, 1.0, 0.0]
    ul, uu = [-5.0, -5.0], [5.0, 5.0]
    p_obs, r_obs = [0.5, 0.5], 0.1

    def stage_con(x, u):
        e = x[0:2] - p_obs
        return vcat(ul - u, u - uu, r_obs ** 2.0 - dot(e, e))  # test/car.jl:45-53

    def term_con(x, u):
        e = x[0:2] - p_obs
        return vcat(x - xT, r_obs ** 2.0 - dot(e, e))  # test/car.jl:54-60

    return Model(
        "car",
        Dynamics(car_discrete, n, m),
        Cost(lambda x, u: 1.0 * dot(x - xT, x - xT) + 1.0e-2 * dot(u, u), n, m),
        Cost(lambda x, u: 1000.0 * dot(x - xT, x - xT), n, 0),
        Constraint(stage_con, n, m, indices_inequality=range(5)),
        Constraint(term_con, n, 0, indices_inequality=[3]),
    )


# ----------------------------------------------------------------------------- pendulum (unit fixture)
def pendulum_continuous(x, u):
    mass, length_com, gravity, damping = 1.0, 1.0, 9.81, 0.1  # test/dynamics.jl:8-15
    return SymVec([x[1],
                   (u[0] / (mass * length_com * length_com)
                    - gravity * sp.sin(x[0]) / length_com
                    - damping * x[1] / (mass * length_com * length_com))])


def pendulum_discrete(x, u):
    h = 0.1  # explicit Euler, test/dynamics.jl:17-19
    return x + h * pendulum_continuous(x, u)


@functools.lru_cache(maxsize=None)
def pendulum() -> Model:
    """Unconstrained swing-up used to exercise the plain (non-AL) solve! path
    (src/solve.jl:137-139) with the unit-test dynamics and test/objective.jl:6-7 costs."""
    n, m = 2, 1
    return Model(
        "pendulum",
        Dynamics(pendulum_discrete, n, m),
        Cost(lambda x, u: dot(x, x) + 0.1 * dot(u, u), n, m),
        Cost(lambda x, u: 10.0 * dot(x, x), n, 0),
    )


# ----------------------------------------------------------------------------- dense LQ tracking (config 4)
@functools.lru_cache(maxsize=None)
def lq_tracking(n: int = 64, m: int = 16, seed: int = 1) -> Model:
    """f = A diag(1 + 0.02 s) x + B u with w = [s; r]; cost 1/2 (x-r)'Q(x-r) + 1/2 u'Ru, Q = I, R = 0.1 I
    (SURVEY.md section 8d, config C4), with A = 0.7 I + 0.1 G / sqrt(n), B = 0.1 G'.

    SURVEY proposes A = I + 0.05 G / sqrt(n) and 1 + 0.05 s.  Measured with the oracle: for ||A||_2 >~ 0.9 the
    reference's un-symmetrised value recursion P = K'QuuK + K'Qux + Qux'K + Qxx (src/backward_pass.jl:79-84)
    loses symmetry exponentially over the 255 steps at n = 64 (|P - P'| ~ 1e16 after 150 steps), Quu stops being
    positive definite, potrf fails silently (Q3) and every line search fails -- in the oracle and, bit for bit,
    in the engine.  The synthetic plant is therefore made contractive (||A||_2 = 0.83), for which the recursion
    stays symmetric to 1e-15."""
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    A = 0.7 * np.eye(n) + 0.1 * G / math.sqrt(n)
    B = 0.1 * rng.standard_normal((n, m))
    p = 2 * n

    def dyn(x, u, w):
        xs = [x[j] * (1.0 + 0.02 * w[j]) for j in range(n)]
        return [sum((float(A[i, j]) * xs[j] for j in range(n)), sp.Integer(0))
                + sum((float(B[i, k]) * u[k] for k in range(m)), sp.Integer(0)) for i in range(n)]

    def stage(x, u, w):
        e = [x[i] - w[n + i] for i in range(n)]
        return 0.5 * dot(e, e) + 0.5 * 0.1 * dot(u, u)

    def term(x, u, w):
        e = [x[i] - w[n + i] for i in range(n)]
        return 0.5 * dot(e, e)

    return Model(f"lq{n}x{m}", Dynamics(dyn, n, m, p), Cost(stage, n, m, p), Cost(term, n, 0, p))


@functools.lru_cache(maxsize=None)
def lq_invariant(n: int = 64, m: int = 16, seed: int = 2) -> Model:
    """A linear time-invariant plant f = A x + B u (A, B as in ``lq_tracking``) tracking a reference that arrives through
    the parameters, w = r: the generated Jacobians contain no x, u, w (ILQR_JAC_CONST), so the wide-model path keeps ONE
    staged Jacobian block for the whole batch."""
    rng = np.random.default_rng(seed)
    A = 0.7 * np.eye(n) + 0.1 * rng.standard_normal((n, n)) / math.sqrt(n)
    B = 0.1 * rng.standard_normal((n, m))

    def dyn(x, u, w):
        return [sum((float(A[i, j]) * x[j] for j in range(n)), sp.Integer(0))
                + sum((float(B[i, k]) * u[k] for k in range(m)), sp.Integer(0)) for i in range(n)]

    def stage(x, u, w):
        e = [x[i] - w[i] for i in range(n)]
        return 0.5 * dot(e, e) + 0.5 * 0.1 * dot(u, u)

    def term(x, u, w):
        e = [x[i] - w[i] for i in range(n)]
        return 0.5 * dot(e, e)

    return Model(f"lqinv{n}x{m}", Dynamics(dyn, n, m, n), Cost(stage, n, m, n), Cost(term, n, 0, n))


@functools.lru_cache(maxsize=None)
def lq_banded(n: int = 64, m: int = 16) -> Model:
    """Same shapes and cost as ``lq_tracking`` but a banded plant (x+_i = 0.7 (1 + 0.02 s_i) x_i + 0.1 x_{i+1}
    + 0.1 u_{i mod m}), so that the generated functions are a few hundred statements and the wide-model kernels
    compile in seconds.  The Riccati kernel's work does not depend on the plant's sparsity (it runs dense
    contractions), which makes this the development / profiling stand-in for BASELINE config 4."""
    p = 2 * n

    def dyn(x, u, w):
        return [0.7 * (1.0 + 0.02 * w[i]) * x[i] + 0.1 * x[(i + 1) % n] + 0.1 * u[i % m] for i in range(n)]

    def stage(x, u, w):
        e = [x[i] - w[n + i] for i in range(n)]
        return 0.5 * dot(e, e) + 0.5 * 0.1 * dot(u, u)

    def term(x, u, w):
        e = [x[i] - w[n + i] for i in range(n)]
        return 0.5 * dot(e, e)

    return Model(f"lqband{n}x{m}", Dynamics(dyn, n, m, p), Cost(stage, n, m, p), Cost(term, n, 0, p))
