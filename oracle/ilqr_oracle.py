"""CPU ORACLE (test infrastructure -- never imported by the product package).

A literal numpy restatement of IterativeLQR.jl's solve path for ONE problem,
array-of-small-arrays like the reference, following its statement order so that
the documented quirks (Hessian accumulation Q1, stale
constraint values Q2, ...) fall out of the code rather than being special-cased.
Each function cites the reference file:line it restates (paths relative to
/root/reference/).

PARITY STATUS: "parity unpinned" against the Julia package itself -- Julia and
Symbolics.jl are absent from this image and the reference's own tests hold no
golden iteration histories.  What pins this file is
tests/test_oracle.py and tests/test_codegen.py (the reference's five test files restated)
and, transitively, oracle/ilqr_oracle.c and the CUDA engine, which must match it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.

Model callables use the reference's convention fn(out, x, u, w) -> None writing
dense arrays in place (src/dynamics.jl:36-37).  ``indices_inequality`` is 0-based.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.linalg.lapack import dpotrf, dpotrs


@dataclass
class Options:
    """src/options.jl:1-14"""
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
    verbose: bool = False


# ----------------------------------------------------------------------------- rollout.jl
def rollout(dynamics, initial_state, actions, parameters=None):
    """src/rollout.jl:33-42 (open loop)."""
    if parameters is None:
        parameters = [np.zeros(d.num_parameter) for d in dynamics]
    x_history = [np.array(initial_state, dtype=float)]
    for t, d in enumerate(dynamics):
        d.evaluate(d.evaluate_cache, x_history[-1], actions[t], parameters[t])  # dynamics! src/dynamics.jl:36-39
        x_history.append(d.evaluate_cache.copy())
    return x_history


class OracleSolver:
    """Solver + ProblemData + PolicyData + SolverData (+ AugmentedLagrangianCosts) in one
    object: src/solver.jl:4-46, src/data/problem.jl:25-45, src/data/policy.jl:44-77,
    src/data/solver.jl:20-47, src/augmented_lagrangian.jl:13-37."""

    def __init__(self, dynamics, objective, constraints=None, parameters=None, options=None):
        self.dynamics = list(dynamics)
        self.costs = list(objective)
        self.constraints = None if constraints is None else list(constraints)
        self.options = options or Options()
        H = len(self.dynamics) + 1
        assert len(self.costs) == H  # src/data/problem.jl:30
        self.H = H
        dyn = self.dynamics
        if parameters is None:
            parameters = [np.zeros(d.num_parameter) for d in dyn] + [np.zeros(0)]
        parameters = [np.asarray(w, dtype=float) for w in parameters]
        if len(parameters) == len(dyn):
            parameters = parameters + [np.zeros(0)]  # src/data/problem.jl:28
        assert len(parameters) == H
        self.parameters = parameters
        ns = [d.num_state for d in dyn] + [dyn[-1].num_next_state]
        ms = [d.num_action for d in dyn]
        self.ns, self.ms = ns, ms
        # current and nominal trajectories (src/data/problem.jl:32-38): zeros at creation
        self.states = [np.zeros(n) for n in ns]
        self.actions = [np.zeros(m) for m in ms] + [np.zeros(0)]
        self.nominal_states = [np.zeros(n) for n in ns]
        self.nominal_actions = [np.zeros(m) for m in ms] + [np.zeros(0)]
        # model data (src/data/model.jl:12-17)
        self.fx = [np.zeros((d.num_next_state, d.num_state)) for d in dyn]
        self.fu = [np.zeros((d.num_next_state, d.num_action)) for d in dyn]
        # objective data (src/data/objective.jl:12-21)
        self.gx = [np.zeros(n) for n in ns]
        self.gu = [np.zeros(m) for m in ms]
        self.gxx = [np.zeros((n, n)) for n in ns]
        self.guu = [np.zeros((m, m)) for m in ms]
        self.gux = [np.zeros((m, n)) for m, n in zip(ms, ns)]
        # policy data (src/data/policy.jl:44-77)
        self.K = [np.zeros((m, n)) for m, n in zip(ms, ns)]
        self.k = [np.zeros(m) for m in ms]
        self.P = [np.zeros((n, n)) for n in ns]
        self.p = [np.zeros(n) for n in ns]
        self.Qx = [np.zeros(n) for n in ns[:-1]]
        self.Qu = [np.zeros(m) for m in ms]
        self.Qxx = [np.zeros((n, n)) for n in ns[:-1]]
        self.Quu = [np.zeros((m, m)) for m in ms]
        self.Qux = [np.zeros((m, n)) for m, n in zip(ms, ns)]
        # solver data (src/data/solver.jl:20-47)
        n_total = sum(ns)
        self.indices_state, self.indices_action = [], []
        n_sum = m_sum = 0
        for t in range(H - 1):
            self.indices_state.append(np.arange(n_sum, n_sum + ns[t]))
            self.indices_action.append(np.arange(n_total + m_sum, n_total + m_sum + ms[t]))
            n_sum += ns[t]
            m_sum += ms[t]
        self.indices_state.append(np.arange(n_sum, n_sum + ns[-1]))
        nz = n_total + sum(ms)
        self.objective = [math.inf]
        self.gradient = np.zeros(nz)
        self.trajectory = np.zeros(nz)
        self.max_violation = [0.0]
        self.step_size = [1.0]
        self.status = [False]
        self.iterations = [0]
        # augmented Lagrangian (src/augmented_lagrangian.jl:13-37, src/data/constraints.jl:11-17)
        if self.constraints is not None:
            assert len(self.constraints) == H
            cons = self.constraints
            self.c = [np.zeros(c.num_constraint) for c in cons]
            self.cx = [np.zeros((c.num_constraint, ns[t])) for t, c in enumerate(cons)]
            self.cu = [np.zeros((c.num_constraint, ms[t])) for t, c in enumerate(cons[:-1])]
            self.rho = [np.ones(c.num_constraint) for c in cons]
            self.lam = [np.zeros(c.num_constraint) for c in cons]
            self.active = [np.ones(c.num_constraint, dtype=int) for c in cons]
        # instrumentation (not in the reference): the per-iteration printout of src/solve.jl:40-45
        self.history = []
        self.chol_fail = False
        self._outer = 0
        # ablation switches for tests (default = reference behaviour)
        self.accumulate_hessians = True   # Q1
        self.stale_constraints = True     # Q2

    # ------------------------------------------------------------------ src/solver.jl:48-66
    def initialize_controls(self, actions):
        for t, ut in enumerate(actions):
            self.nominal_actions[t][...] = ut

    def initialize_states(self, states):
        for t, xt in enumerate(states):
            self.nominal_states[t][...] = xt

    def get_trajectory(self):
        return self.nominal_states, self.nominal_actions[:-1]

    def current_trajectory(self):
        return self.states, self.actions[:-1]

    # ------------------------------------------------------------------ helpers
    @property
    def constrained(self):
        return self.constraints is not None

    def _trajectories(self, mode):
        # src/data/methods.jl:56-61
        if mode == "nominal":
            return self.nominal_states, self.nominal_actions, self.parameters
        return self.states, self.actions, self.parameters

    # ------------------------------------------------------------------ costs
    def _plain_cost(self, states, actions, parameters):
        """src/costs.jl:48-55"""
        J = 0.0
        for t, cost in enumerate(self.costs):
            cost.evaluate(cost.evaluate_cache, states[t], actions[t], parameters[t])
            J += cost.evaluate_cache[0]
        return J

    def _constraint(self, states, actions, parameters):
        """constraint! src/constraints.jl:66-73 via src/data/constraints.jl:19-21"""
        for t, con in enumerate(self.constraints):
            if con.num_constraint == 0:
                continue
            con.evaluate(con.evaluate_cache, states[t], actions[t], parameters[t])
            self.c[t][...] = con.evaluate_cache
            con.evaluate_cache[...] = 0.0

    def _active_set(self):
        """src/augmented_lagrangian.jl:68-85"""
        for t, con in enumerate(self.constraints):
            self.active[t][...] = 1
            for i in con.indices_inequality:
                if self.c[t][i] < 0.0 and self.lam[t][i] == 0.0:
                    self.active[t][i] = 0

    def _al_cost(self, states, actions, parameters):
        """src/augmented_lagrangian.jl:39-66"""
        J = self._plain_cost(states, actions, parameters)
        self._constraint(states, actions, parameters)
        self._active_set()
        for t in range(self.H):
            J += float(self.lam[t] @ self.c[t]) if len(self.c[t]) else 0.0
            for i in range(self.constraints[t].num_constraint):
                if self.active[t][i] == 1:
                    J += 0.5 * self.rho[t][i] * self.c[t][i] ** 2.0
        return J

    def _constraint_violation(self, states, actions, parameters):
        """src/data/constraints.jl:23-46 (always the inf-norm, Q8)"""
        self._constraint(states, actions, parameters)
        return self._violation_of_buffer()

    def _violation_of_buffer(self):
        max_violation = 0.0
        for t, con in enumerate(self.constraints):
            ineq = con.indices_inequality
            for i in range(con.num_constraint):
                c = self.c[t][i]
                cti = max(0.0, c) if i in ineq else abs(c)
                max_violation = max(max_violation, cti)
        return max_violation

    def cost_bang(self, mode):
        """cost! src/data/methods.jl:13-30"""
        x, u, w = self._trajectories(mode)
        if self.constrained:
            self.objective[0] = self._al_cost(x, u, w)
            if self.stale_constraints:
                # Q2: ALWAYS on the current trajectory; overwrites c (not the active set)
                self.max_violation[0] = self._constraint_violation(self.states, self.actions, self.parameters)
            else:
                self.max_violation[0] = self._violation_of_buffer()
        else:
            self.objective[0] = self._plain_cost(x, u, w)
        return self.objective

    # ------------------------------------------------------------------ gradients.jl
    def gradients(self, mode="nominal"):
        """gradients! src/gradients.jl:92-98 -> :1-8, :10-21, :23-81, :83-90"""
        x, u, w = self._trajectories(mode)
        H = self.H
        # dynamics Jacobians, overwrite (src/dynamics.jl:41-50)
        for t, d in enumerate(self.dynamics):
            d.jacobian_state(d.jacobian_state_cache, x[t], u[t], w[t])
            d.jacobian_action(d.jacobian_action_cache, x[t], u[t], w[t])
            self.fx[t][...] = d.jacobian_state_cache
            self.fu[t][...] = d.jacobian_action_cache
            d.jacobian_state_cache[...] = 0.0
            d.jacobian_action_cache[...] = 0.0
        # cost gradients, overwrite (src/costs.jl:57-68)
        for t, cost in enumerate(self.costs):
            cost.gradient_state(cost.gradient_state_cache, x[t], u[t], w[t])
            self.gx[t][...] = cost.gradient_state_cache
            cost.gradient_state_cache[...] = 0.0
            if t == H - 1:
                continue
            cost.gradient_action(cost.gradient_action_cache, x[t], u[t], w[t])
            self.gu[t][...] = cost.gradient_action_cache
            cost.gradient_action_cache[...] = 0.0
        # cost Hessians, ACCUMULATE (src/costs.jl:70-84) -- Q1
        for t, cost in enumerate(self.costs):
            if not self.accumulate_hessians:
                self.gxx[t][...] = 0.0
                if t < H - 1:
                    self.guu[t][...] = 0.0
                    self.gux[t][...] = 0.0
            cost.hessian_state_state(cost.hessian_state_state_cache, x[t], u[t], w[t])
            self.gxx[t] += cost.hessian_state_state_cache
            cost.hessian_state_state_cache[...] = 0.0
            if t == H - 1:
                continue
            cost.hessian_action_action(cost.hessian_action_action_cache, x[t], u[t], w[t])
            cost.hessian_action_state(cost.hessian_action_state_cache, x[t], u[t], w[t])
            self.guu[t] += cost.hessian_action_action_cache
            self.gux[t] += cost.hessian_action_state_cache
            cost.hessian_action_action_cache[...] = 0.0
            cost.hessian_action_state_cache[...] = 0.0
        if not self.constrained:
            return
        # constraint Jacobians, overwrite (src/constraints.jl:75-87)
        for t, con in enumerate(self.constraints):
            if con.num_constraint == 0:
                continue
            con.jacobian_state(con.jacobian_state_cache, x[t], u[t], w[t])
            self.cx[t][...] = con.jacobian_state_cache
            con.jacobian_state_cache[...] = 0.0
            if t == H - 1:
                continue
            con.jacobian_action(con.jacobian_action_cache, x[t], u[t], w[t])
            self.cu[t][...] = con.jacobian_action_cache
            con.jacobian_action_cache[...] = 0.0
        # AL terms (src/gradients.jl:54-80); c and a are whatever cost! left behind (Q2)
        for t in range(H):
            d = self.rho[t] * self.active[t]                 # Irho diagonal, :56-58
            c_tmp = self.lam[t] + d * self.c[t]              # :59-62
            self.gx[t] += self.cx[t].T @ c_tmp               # :63
            cx_tmp = d[:, None] * self.cx[t]                 # :66
            self.gxx[t] += self.cx[t].T @ cx_tmp             # :67
            if t == H - 1:
                continue
            self.gu[t] += self.cu[t].T @ c_tmp               # :72
            cu_tmp = d[:, None] * self.cu[t]                 # :75
            self.guu[t] += self.cu[t].T @ cu_tmp             # :76
            self.gux[t] += self.cu[t].T @ cx_tmp             # :79

    # ------------------------------------------------------------------ backward_pass.jl
    def backward_pass(self):
        """src/backward_pass.jl:39-90"""
        H = self.H
        fx, fu, gx, gu, gxx, guu, gux = self.fx, self.fu, self.gx, self.gu, self.gxx, self.guu, self.gux
        P, p, K, k = self.P, self.p, self.K, self.k
        P[H - 1][...] = gxx[H - 1]
        p[H - 1][...] = gx[H - 1]
        for t in range(H - 2, -1, -1):
            self.Qx[t][...] = fx[t].T @ p[t + 1] + gx[t]                 # :44-45
            self.Qu[t][...] = fu[t].T @ p[t + 1] + gu[t]                 # :48-49
            self.Qxx[t][...] = (fx[t].T @ P[t + 1]) @ fx[t] + gxx[t]     # :52-54
            uxh = fu[t].T @ P[t + 1]                                      # :57 / :62
            self.Quu[t][...] = uxh @ fu[t] + guu[t]                      # :58-59
            self.Qux[t][...] = uxh @ fx[t] + gux[t]                      # :63-64
            # Cholesky, upper, no regularisation, info ignored (:68-75, Q3)
            U, info = dpotrf(self.Quu[t], lower=0, clean=0, overwrite_a=0)
            if info != 0:
                self.chol_fail = True
            Kt, _ = dpotrs(U, self.Qux[t], lower=0)
            kt, _ = dpotrs(U, self.Qu[t], lower=0)
            K[t][...] = -1.0 * Kt.reshape(K[t].shape)
            k[t][...] = -1.0 * kt.reshape(k[t].shape)
            ux_tmp = self.Quu[t] @ K[t]                                   # :79
            Pt = K[t].T @ ux_tmp                                          # :81
            Pt = Pt + K[t].T @ self.Qux[t]                                # :82
            Pt = Pt + self.Qux[t].T @ K[t]                                # :83
            P[t][...] = Pt + self.Qxx[t]                                  # :84
            pt = ux_tmp.T @ k[t]                                          # :86
            pt = pt + K[t].T @ self.Qu[t]                                 # :87
            pt = pt + self.Qux[t].T @ k[t]                                # :88
            p[t][...] = pt + self.Qx[t]                                   # :89

    # ------------------------------------------------------------------ solve.jl:67-83
    def lagrangian_gradient(self):
        for t in range(self.H - 1):
            self.gradient[self.indices_state[t]] = self.Qx[t] - self.p[t]
            self.gradient[self.indices_action[t]] = self.Qu[t]

    # ------------------------------------------------------------------ data/methods.jl
    def trajectory_sensitivities(self):
        """src/data/methods.jl:42-54"""
        self.trajectory[...] = 0.0
        for t in range(self.H - 1):
            zx = self.trajectory[self.indices_state[t]]
            zu = self.k[t] + self.K[t] @ zx
            self.trajectory[self.indices_action[t]] = zu
            self.trajectory[self.indices_state[t + 1]] = self.fu[t] @ zu + self.fx[t] @ zx

    def update_nominal_trajectory(self):
        """src/data/methods.jl:32-39"""
        for t in range(self.H):
            self.nominal_states[t][...] = self.states[t]
            if t == self.H - 1:
                continue
            self.nominal_actions[t][...] = self.actions[t]

    # ------------------------------------------------------------------ rollout.jl:1-31
    def rollout_bang(self, step_size=1.0):
        x, u, w = self.states, self.actions, self.parameters
        xb, ub = self.nominal_states, self.nominal_actions
        x[0][...] = xb[0]
        for t, d in enumerate(self.dynamics):
            u[t][...] = self.k[t]
            u[t] *= step_size
            u[t] += ub[t]
            u[t] += self.K[t] @ x[t]        # mul!(u, K, x, 1, 1)
            u[t] += -1.0 * (self.K[t] @ xb[t])  # mul!(u, K, xbar, -1, 1)
            d.evaluate(d.evaluate_cache, x[t], u[t], w[t])
            x[t + 1][...] = d.evaluate_cache

    # ------------------------------------------------------------------ forward_pass.jl
    def forward_pass(self, c1=1.0e-4, max_iterations=25):
        """src/forward_pass.jl:1-56"""
        opts = self.options
        self.status[0] = False
        J_prev = self.objective[0]
        self.lagrangian_gradient()
        if opts.line_search == "armijo":
            self.trajectory_sensitivities()
            delta_grad_product = float(self.gradient @ self.trajectory)
        else:
            delta_grad_product = 0.0
        self.step_size[0] = 1.0
        iteration = 1
        while self.step_size[0] >= opts.min_step_size:
            if iteration > max_iterations:
                break
            self.rollout_bang(step_size=self.step_size[0])
            J = self.cost_bang("current")[0]
            if J <= J_prev + c1 * self.step_size[0] * delta_grad_product:
                self.update_nominal_trajectory()
                self.objective[0] = J
                self.status[0] = True
                break
            else:
                self.step_size[0] *= 0.5
                iteration += 1

    # ------------------------------------------------------------------ solve.jl:1-54
    def _reset_model_objective(self):
        """reset!(problem.model); reset!(problem.objective)  src/solve.jl:9-10"""
        for a in self.fx + self.fu + self.gx + self.gu + self.gxx + self.guu + self.gux:
            a[...] = 0.0

    def _reset_data(self):
        """reset!(data) src/data/solver.jl:49-59"""
        self.objective[0] = 0.0
        self.gradient[...] = 0.0
        self.max_violation[0] = 0.0
        self.status[0] = False
        self.iterations[0] = 0

    def ilqr_solve(self):
        opts = self.options
        self._reset_model_objective()
        if opts.reset_cache:
            self._reset_data()
        self.cost_bang("nominal")
        self.gradients("nominal")
        self.backward_pass()
        obj_prev = self.objective[0]
        for i in range(1, opts.max_iterations + 1):
            self.forward_pass()
            if opts.line_search != "none":
                self.gradients("nominal")
                self.backward_pass()
                self.lagrangian_gradient()
            gradient_norm = float(np.max(np.abs(self.gradient))) if self.gradient.size else 0.0
            self.iterations[0] += 1
            self.history.append(dict(outer=self._outer, iter=i, cost=self.objective[0],
                                     gradient_norm=gradient_norm, max_violation=self.max_violation[0],
                                     step_size=self.step_size[0], status=self.status[0]))
            if gradient_norm < opts.lagrangian_gradient_tolerance:
                break
            if abs(self.objective[0] - obj_prev) < opts.objective_tolerance:
                break
            else:
                obj_prev = self.objective[0]
            if not self.status[0]:
                break

    def constrained_ilqr_solve(self, augmented_lagrangian_callback=None):
        """src/solve.jl:88-129"""
        opts = self.options
        self._reset_data()
        for lam in self.lam:
            lam[...] = 0.0
        for rho in self.rho:
            rho[...] = opts.initial_constraint_penalty
        for i in range(1, opts.max_dual_updates + 1):
            self._outer = i
            self.ilqr_solve()
            self.cost_bang("nominal")
            if self.max_violation[0] <= opts.constraint_tolerance:
                break
            self.augmented_lagrangian_update(opts.scaling_penalty, opts.max_penalty)
            if augmented_lagrangian_callback is not None:
                augmented_lagrangian_callback(self)

    def augmented_lagrangian_update(self, scaling_penalty=10.0, max_penalty=1.0e12):
        """src/augmented_lagrangian.jl:87-110"""
        for t, con in enumerate(self.constraints):
            for i in range(con.num_constraint):
                self.lam[t][i] += self.rho[t][i] * self.c[t][i]
                if i in con.indices_inequality:
                    self.lam[t][i] = max(0.0, self.lam[t][i])
                self.rho[t][i] = min(scaling_penalty * self.rho[t][i], max_penalty)

    def solve(self, states=None, actions=None, **kwargs):
        """solve! src/solve.jl:137-143 (+ warm-start forms :56-60, :131-135)"""
        if actions is not None:
            self.initialize_controls(actions)
        if states is not None:
            self.initialize_states(states)
        if self.constrained:
            self.constrained_ilqr_solve(**kwargs)
        else:
            self._outer = 0
            self.ilqr_solve()
