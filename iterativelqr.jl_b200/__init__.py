"""B200-native batched iLQR engine: host-side mirror of IterativeLQR.jl's API.

Names follow /root/reference/src/IterativeLQR.jl:31-45 (``!`` dropped): Dynamics, Cost,
Constraint, Solver, Options, rollout, initialize_controls, initialize_states, solve,
get_trajectory, current_trajectory.  The numerical work is done by libilqr_cuda.so
(include/ilqr_cuda.h) -- CUDA kernels for sm_100a; there is no CPU fallback.
"""
from .api import Constraint, Cost, Dynamics, Model, constraint_from_c, dynamics_from_c  # noqa: F401
from .solver import (Options, Solver, current_trajectory, get_trajectory, initialize_controls,  # noqa: F401
                     initialize_states, rollout, solve, solve_stream)
from .codegen import dot, vcat  # noqa: F401
