"""mpc_fatigue_b200 — B200-native batched evaluator for the hot path of ADVRHumanoids/mpc_fatigue:
rigid-body dynamics (RNEA / forward dynamics, frame FK, frame Jacobians) + the joint fatigue/thermal
compartment ODE, RK4 over every (scenario, node) unit, with forward-mode Jacobians.

Importing the package loads the in-tree CUDA library (lib/libmpcf.so) and fails loudly if it is missing:
there is no CPU fallback.
"""
from . import _capi  # noqa: F401  (raises ImportError when libmpcf.so has not been built)
from .model import Model, data_urdf  # noqa: F401

__all__ = ["Model", "data_urdf"]
