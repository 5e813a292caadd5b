#!/usr/bin/env python3
"""Replay a stored solution of the reference's dual-arm box OCP through the GPU evaluator.

The reference replays its IPOPT solution with one CasADi call per node and per arm
(python/2_pilz_6_DOF/Box_Pilz_6DOF2.py:534-546: `Idyn_LR(q=..., qdot=..., qddot=...)['tau'] - mtimes(J_LR.T, W_LR)`)
and then plots the torques.  Here every node of both arms is evaluated in a handful of launches, first through the
reference-style Function objects (drop-in call convention), then through the batched OCP layer.

    python examples/replay_box_solution.py [Result_1|Result_2|Result_4|plotter]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mpc_fatigue_b200.pynocchio_casadi as pin  # noqa: E402  (was: import mpc_fatigue.pynocchio_casadi as pin)
from mpc_fatigue_b200.model import data_urdf  # noqa: E402
from mpc_fatigue_b200.ocp import DualArmBoxOCP  # noqa: E402
from mpc_fatigue_b200.pynocchio_casadi import Function  # noqa: E402  (was: from casadi import *)

key = sys.argv[1] if len(sys.argv) > 1 else "Result_2"
sol = DualArmBoxOCP.parse_solution(np.load(os.path.join(ROOT, "tests", "golden", "plotter_solutions.npz"))[key])
N = sol["N"]
urdf_1, urdf_2 = data_urdf("pilz6_first"), data_urdf("pilz6_second")

# ---- the reference idiom, all N nodes per call instead of a Python loop over k ----
Idyn_RR = Function.deserialize(pin.generate_inv_dyn(urdf_2))
jac_RR = Function.deserialize(pin.generate_jacobian(urdf_2, "end_effector"))
q_RR, qd_RR = sol["q"][:N, 6:], sol["qd"][:, 6:]
W_RR = np.hstack([sol["F_RR"], np.zeros((N, 3))])
J = jac_RR(q=q_RR)["J"]                                   # [N, 6, 6]
tau_RR = Idyn_RR(q=q_RR, qdot=qd_RR, qddot=0 * q_RR)["tau"] - np.einsum("kri,kr->ki", J, W_RR)

# ---- the batched OCP layer: both arms, all constraint rows ----
ocp = DualArmBoxOCP(urdf_1, urdf_2)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda().unsqueeze(0)
rows = ocp.evaluate(t(sol["q"]), t(sol["qd"]), t(sol["F_LR"]), t(sol["F_RR"]))
assert np.abs(rows["tau_R"][0].cpu().numpy() - tau_RR).max() < 1e-9

np.set_printoptions(precision=4, suppress=True, linewidth=160)
print("%s: N = %d nodes, h = %.4f s" % (key, N, 2.0 / N))
print("max Euler defect      L %.2e   R %.2e" % (float(rows["defect_L"].abs().max()), float(rows["defect_R"].abs().max())))
print("max |dist^2 - 0.04|   %.2e" % float((rows["dist2"] - 0.04).abs().max()))
print("max force / moment equilibrium residual  %.2e / %.2e" % (float(rows["force_eq"].abs().max()), float(rows["moment_eq"].abs().max())))
print("right-arm torques, last third of the horizon (bounds of Box_Pilz_6DOF.py: tau_0,1 in [-5, 5], tau_2 in [-10, 5]):")
print(rows["tau_R"][0, 2 * N // 3::4].cpu().numpy())
