"""Synthetic scenario batches (SURVEY.md §8d value distributions).

Every value is a pure function of (seed, global scenario index, node, channel, joint) through a
counter-based hash (splitmix64 finaliser on int64 tensors), so a scenario's inputs do not depend on how
the batch is sharded across ranks: results for scenario i are world-size invariant.

Unit order is node-major inside a shard: u = k * B_local + b (b = local scenario), which keeps the N
nodes of one scenario on one GPU and makes the per-scenario reduction coalesced.
"""
from __future__ import annotations

import torch

_K1 = 0xBF58476D1CE4E5B9 - (1 << 64)
_K2 = 0x94D049BB133111EB - (1 << 64)
_GOLD = 0x9E3779B97F4A7C15 - (1 << 64)


def _mix(x: torch.Tensor) -> torch.Tensor:
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * _K1
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * _K2
    return x ^ ((x >> 31) & ((1 << 33) - 1))


def uniform01(seed: int, scenario: torch.Tensor, node: torch.Tensor, channel: int, joint: torch.Tensor) -> torch.Tensor:
    """U[0,1) for broadcastable int64 index tensors."""
    ctr = ((scenario * 1000003 + node) * 64 + joint) * 8 + channel
    x = _mix(ctr * _GOLD + int(seed))
    return ((x >> 11) & ((1 << 53) - 1)).to(torch.float64) * (1.0 / 9007199254740992.0)


def synth_batch(limits: dict, scenario_start: int, B: int, N: int, seed: int = 1234, device="cuda"):
    """Return q, qd, tau, f as [n, N*B] float64 tensors (node-major units) for scenarios
    [scenario_start, scenario_start + B).  limits: q_lo, q_hi, v_max, tau_max arrays of length n."""
    dev = torch.device(device)
    n = len(limits["q_lo"])
    t = lambda k: torch.as_tensor(limits[k], dtype=torch.float64, device=dev).reshape(n, 1)
    q_lo, q_hi, v_max, tau_max = t("q_lo"), t("q_hi"), t("v_max"), t("tau_max")
    scen = (torch.arange(B, dtype=torch.int64, device=dev) + int(scenario_start)).reshape(1, 1, B)
    node = torch.arange(N, dtype=torch.int64, device=dev).reshape(1, N, 1)
    joint = torch.arange(n, dtype=torch.int64, device=dev).reshape(n, 1, 1)
    U = N * B
    u = lambda ch: uniform01(seed, scen, node, ch, joint).reshape(n, U)
    q = q_lo + (q_hi - q_lo) * u(0)
    qd = (2.0 * u(1) - 1.0) * v_max
    tau = (2.0 * u(2) - 1.0) * (0.25 * tau_max)
    f = 20.0 + 60.0 * u(3)
    return q.contiguous(), qd.contiguous(), tau.contiguous(), f.contiguous()
