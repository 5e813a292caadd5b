"""Receding-horizon plumbing around the batched node evaluator.

Follows python/Centauro_script/mpc_principal.py (solution layout :120-129, next initial condition :365-374, the
per-node post-solve loop :395-415) and python/Centauro_script/unroller_node.py:168-190 (linear interpolation between
nodes).  The reference walks the N nodes one CasADi call at a time; here every node of every scenario goes through
one batched launch.  IPOPT / CasADi are not part of this repository: `RecedingHorizonDriver` takes the NLP solve as a
callable.
"""
from __future__ import annotations

import numpy as np
import torch

from .evaluator import BatchEvaluator
from .model import Model


class ThermalSolutionLayout:
    """Decision-vector layout of the thermal MPC (mpc_principal.py:120-129): per node [q(nq), T(nq), qd(nq), F(2 nf)],
    then the terminal [q_N(nq), T_N(nq)]; length N (3 nq + 2 nf) + 2 nq."""

    def __init__(self, nq: int, nf: int = 3):
        self.nq, self.nf = nq, nf
        self.stride = 3 * nq + 2 * nf

    def horizon(self, length: int) -> int:
        N, rem = divmod(length - 2 * self.nq, self.stride)
        if N < 1 or rem:
            raise ValueError("vector length %d is not N*%d+%d" % (length, self.stride, 2 * self.nq))
        return N

    def parse(self, vec) -> dict:
        """vec [..., N*stride + 2 nq] -> q [..., N+1, nq], T [..., N+1, nq], qd [..., N, nq], F [..., N, 2 nf]."""
        v = np.asarray(vec, dtype=np.float64)
        N, nq = self.horizon(v.shape[-1]), self.nq
        lead = v.shape[:-1]
        nodes = v[..., :N * self.stride].reshape(lead + (N, self.stride))
        tail = v[..., N * self.stride:]
        cat = lambda a, b: np.concatenate([a, b[..., None, :]], axis=-2)
        return {"N": N, "q": cat(nodes[..., :nq], tail[..., :nq]), "T": cat(nodes[..., nq:2 * nq], tail[..., nq:]),
                "qd": nodes[..., 2 * nq:3 * nq], "F": nodes[..., 3 * nq:]}

    def pack(self, q, T, qd, F) -> np.ndarray:
        q, T, qd, F = (np.asarray(a, dtype=np.float64) for a in (q, T, qd, F))
        N = qd.shape[-2]
        if q.shape[-2] != N + 1 or T.shape[-2] != N + 1 or F.shape[-2] != N or F.shape[-1] != 2 * self.nf:
            raise ValueError("q and T need N+1 nodes, qd and F need N")
        nodes = np.concatenate([q[..., :N, :], T[..., :N, :], qd, F], axis=-1)
        return np.concatenate([nodes.reshape(nodes.shape[:-2] + (-1,)), q[..., N, :], T[..., N, :]], axis=-1)

    def warm_start(self, q0, T0, qd0, F0, N: int) -> np.ndarray:
        """First guess of mpc_principal.py:120-121: the initial node repeated over the horizon."""
        node = np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1) for a in (q0, T0, qd0, F0)])
        if len(node) != self.stride:
            raise ValueError("initial node has %d entries, layout needs %d" % (len(node), self.stride))
        return np.concatenate([np.tile(node, N), node[:2 * self.nq]])

    def next_initial_condition(self, sol, temp_offset: float = 0.05, decimals: int = 4) -> dict:
        """Initial condition of the next OCP (mpc_principal.py:365-374): terminal q, terminal T minus 0.05 K, the
        velocity of the last interval, each rounded to 4 decimals.  Batched over leading dimensions."""
        s = self.parse(sol)
        return {"q": np.round(s["q"][..., -1, :], decimals), "T": np.round(s["T"][..., -1, :] - temp_offset, decimals),
                "qd": np.round(s["qd"][..., -1, :], decimals)}


def resample_trajectory(knots: torch.Tensor, h: float, t: torch.Tensor) -> torch.Tensor:
    """Piecewise-linear read-out of node trajectories, the unroller's p0 + (p1 - p0) / h * (t - t0)
    (unroller_node.py:186-190), for every scenario and every query time at once.

    knots [B, N+1, d] (node k at time k h), t [M] or [B, M] seconds -> [B, M, d].  Times outside [0, N h] are clamped
    (the unroller stops when the trajectory is exhausted)."""
    if knots.dim() != 3:
        raise ValueError("knots must be [B, N+1, d]")
    B, K, d = knots.shape
    if K < 2 or h <= 0.0:
        raise ValueError("need at least two nodes and h > 0")
    t = t.to(knots.dtype).to(knots.device)
    if t.dim() == 1:
        t = t.unsqueeze(0).expand(B, -1)
    s = (t / h).clamp(0.0, float(K - 1))
    k0 = s.floor().clamp_max(K - 2).long()
    frac = (s - k0.to(s.dtype)).unsqueeze(-1)
    idx = k0.unsqueeze(-1).expand(-1, -1, d)
    p0, p1 = knots.gather(1, idx), knots.gather(1, idx + 1)
    return p0 + (p1 - p0) * frac


class RecedingHorizonDriver:
    """The outer loop of mpc_principal.py:130-418 for a batch of scenarios.

    `solve(x0 [B, len], ic dict) -> sol [B, len]` stands for the NLP solve (IPOPT in the reference).  After each solve the
    driver (a) turns the world-frame contact forces of every node into end-effector axes with one batched FK launch per
    arm (the reference loops `ForwKinLA(q_j, "rot")` over the nodes, :395-411), (b) evaluates the torque / thermal rows of
    every node in one launch so the caller can log bound violations, and (c) forms the next initial condition."""

    def __init__(self, model: Model, left_frame: str, right_frame: str, N: int = 40, T: float = 20.0, device=None):
        from .ocp import ThermalMPCNodes
        self.model, self.N, self.h = model, N, T / N
        self.nodes = ThermalMPCNodes(model, [left_frame, right_frame], T=T, N=N, device=device)
        self.ev: BatchEvaluator = self.nodes.ev
        self.layout = ThermalSolutionLayout(model.n, 3)
        self.device = self.ev.device

    def unroll_messages(self, sol) -> dict:
        """sol [B, len] -> per node q [B, N, nq] and the contact forces in end-effector axes F_local [B, N, 6]
        (what mpc_principal.py:411 publishes; the reference transforms the right-arm force with the LEFT rotation,
        :409 — here each force uses its own arm's rotation)."""
        s = self.layout.parse(np.asarray(sol, dtype=np.float64))
        N = s["N"]
        q = torch.from_numpy(np.ascontiguousarray(s["q"][:, :N])).to(self.device)
        F = torch.from_numpy(np.ascontiguousarray(s["F"])).to(self.device)
        B = q.shape[0]
        qs = q.permute(2, 1, 0).reshape(self.model.n, N * B).contiguous()
        out = []
        for e, fr in enumerate(self.nodes.frames):
            _, rot = self.ev.fk(fr, qs)
            R = rot.reshape(3, 3, N, B).permute(3, 2, 0, 1)  # [B, N, 3, 3] world <- end effector
            out.append(torch.einsum("bnji,bnj->bni", R, F[..., 3 * e:3 * e + 3]))  # R^T F
        return {"q": q, "F_local": torch.cat(out, dim=-1)}

    def node_rows(self, sol) -> dict:
        s = self.layout.parse(np.asarray(sol, dtype=np.float64))
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        F = s["F"]
        z = np.zeros(F.shape[:-1] + (3,))
        W = np.concatenate([F[..., :3], z, F[..., 3:], z], axis=-1)
        return self.nodes.evaluate(dev(s["q"]), dev(s["qd"]), dev(W), dev(s["T"]))

    def run(self, solve, q0, T0, qd0, F0, cycles: int):
        """Runs `cycles` OCPs back to back; q0/T0/qd0 [B, nq], F0 [B, 6].  Returns the list of per-cycle records."""
        q0, T0, qd0, F0 = (np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in (q0, T0, qd0, F0))
        x0 = np.stack([self.layout.warm_start(q0[b], T0[b], qd0[b], F0[b], self.N) for b in range(q0.shape[0])])
        ic = {"q": q0, "T": T0, "qd": qd0}
        log = []
        for _ in range(cycles):
            sol = np.asarray(solve(x0, ic), dtype=np.float64)
            if sol.shape != x0.shape:
                raise ValueError("solve() returned shape %s, expected %s" % (sol.shape, x0.shape))
            rows = self.node_rows(sol)
            log.append({"sol": sol, "messages": self.unroll_messages(sol), "rows": rows, "ic": ic})
            ic = self.layout.next_initial_condition(sol)
            x0 = sol  # the reference warm-starts from the previous solution unshifted (mpc_principal.py:359-362)
        return log
