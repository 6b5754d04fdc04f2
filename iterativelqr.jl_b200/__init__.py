"""B200-native batched iLQR engine: host-side mirror of IterativeLQR.jl's API."""
from .api import Constraint, Cost, Dynamics, Model  # noqa: F401
