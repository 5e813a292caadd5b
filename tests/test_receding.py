"""Receding-horizon plumbing: thermal-MPC solution layout, next initial condition, trajectory
resampling on the CPU; the batched post-solve evaluation against the oracle on the GPU."""
import numpy as np
import pytest
import torch

from mpc_fatigue_b200.receding import ThermalSolutionLayout, resample_trajectory


def test_thermal_layout_round_trip_and_reference_indexing():
    nq, nf, N = 14, 3, 5
    lay = ThermalSolutionLayout(nq, nf)
    n = 3 * nq + 2 * nf  # mpc_principal.py:128-129
    v = np.arange(N * n + 2 * nq, dtype=float)
    s = lay.parse(v)
    assert s["N"] == N and s["q"].shape == (N + 1, nq) and s["F"].shape == (N, 6)
    j = 3
    assert np.array_equal(s["q"][j], v[j * n:j * n + nq])  # sol[j*n : j*n+nq], :403
    assert np.array_equal(s["F"][j], v[j * n + 3 * nq:j * n + n])  # :398
    assert np.array_equal(s["q"][N], v[N * n:N * n + nq]) and np.array_equal(s["T"][N], v[N * n + nq:N * n + 2 * nq])  # :365-366
    assert np.array_equal(lay.pack(s["q"], s["T"], s["qd"], s["F"]), v)
    batch = np.stack([v, 2 * v])
    assert np.array_equal(lay.pack(**{k: lay.parse(batch)[k] for k in ("q", "T", "qd", "F")}), batch)
    with pytest.raises(ValueError):
        lay.parse(v[:-1])


def test_warm_start_and_next_initial_condition():
    nq, N = 4, 3
    lay = ThermalSolutionLayout(nq)
    q0, T0, qd0, F0 = np.arange(4.0), 20.0 + np.arange(4.0), np.zeros(4), np.array([0, 0, 24.5, 0, 0, 24.5])
    x0 = lay.warm_start(q0, T0, qd0, F0, N)  # sol0 = node * N + q0 + T0, :120-121
    s = lay.parse(x0)
    assert x0.shape == (N * lay.stride + 2 * nq,) and np.array_equal(s["q"], np.tile(q0, (N + 1, 1))) and np.array_equal(s["F"][2], F0)
    rng = np.random.default_rng(0)
    sol = rng.normal(size=(2, len(x0)))
    ic = lay.next_initial_condition(sol)
    n = lay.stride
    assert np.array_equal(ic["q"], np.round(sol[:, N * n:N * n + nq], 4))
    assert np.array_equal(ic["T"], np.round(sol[:, N * n + nq:N * n + 2 * nq] - 0.05, 4))
    assert np.array_equal(ic["qd"], np.round(sol[:, (N - 1) * n + 2 * nq:(N - 1) * n + 3 * nq], 4))  # k = N-1 after the loop, :367
    with pytest.raises(ValueError):
        lay.warm_start(q0, T0, qd0, F0[:5], N)


def test_resample_matches_the_unroller_formula():
    B, N, d, h = 3, 6, 5, 0.5
    g = torch.Generator().manual_seed(3)
    knots = torch.randn(B, N + 1, d, dtype=torch.float64, generator=g)
    t = torch.tensor([0.0, 0.1, 0.5, 1.26, 2.999, 3.0, 7.0, -1.0], dtype=torch.float64)
    got = resample_trajectory(knots, h, t)
    for m, tn in enumerate(t.tolist()):
        tc = min(max(tn, 0.0), N * h)
        k = min(int(tc / h), N - 1)
        ref = knots[:, k] + (knots[:, k + 1] - knots[:, k]) / h * (tc - k * h)  # unroller_node.py:186
        assert torch.allclose(got[:, m], ref, rtol=0, atol=1e-13)
    assert torch.allclose(resample_trajectory(knots, h, torch.arange(N + 1, dtype=torch.float64) * h), knots, rtol=0, atol=1e-15)
    per_b = resample_trajectory(knots, h, t.unsqueeze(0).expand(B, -1).contiguous())
    assert torch.equal(per_b, got)
    with pytest.raises(ValueError):
        resample_trajectory(knots[:, :1], h, t)


@pytest.mark.gpu
def test_driver_post_solve_rows_match_the_oracle():
    from conftest import oracle_model_from_export
    from mpc_fatigue_b200.model import Model
    from mpc_fatigue_b200.receding import RecedingHorizonDriver
    from oracle.pyoracle import Oracle
    m = Model.synthetic("dual_arm", 14, seed=4, armature=1e-2)
    orc = Oracle(oracle_model_from_export(m))
    N, B, nq = 6, 3, 14
    drv = RecedingHorizonDriver(m, "left_ee", "right_ee", N=N, T=3.0)
    lay = drv.layout
    rng = np.random.default_rng(5)
    calls = []

    def solve(x0, ic):  # stands in for IPOPT: a random feasible-looking vector that keeps the initial condition
        calls.append((x0.copy(), ic))
        s = lay.parse(x0)
        q = s["q"] + 0.3 * rng.normal(size=s["q"].shape)
        q[:, 0] = ic["q"]
        return lay.pack(q, s["T"] + rng.uniform(0, 1, size=s["T"].shape), 0.4 * rng.normal(size=s["qd"].shape),
                        10.0 * rng.normal(size=s["F"].shape))

    q0 = rng.uniform(-1, 1, size=(B, nq))
    log = drv.run(solve, q0, np.full((B, nq), 20.0), np.zeros((B, nq)), np.tile([0, 0, 24.5, 0, 0, 24.5], (B, 1)), cycles=2)
    assert len(log) == 2 and np.array_equal(calls[1][0], log[0]["sol"])  # warm start = previous solution
    assert np.array_equal(calls[1][1]["q"], np.round(lay.parse(log[0]["sol"])["q"][:, -1], 4))
    rec = log[1]
    s = lay.parse(rec["sol"])
    fl, fr = m.frame_id("left_ee"), m.frame_id("right_ee")
    qn = np.ascontiguousarray(s["q"][:, :N].reshape(B * N, nq).T)  # [nq, U], u = b * N + k
    (pL, RL), (pR, RR) = orc.fk(fl, qn), orc.fk(fr, qn)
    RL, RR = RL.T.reshape(B, N, 3, 3), RR.T.reshape(B, N, 3, 3)
    Floc = np.concatenate([np.einsum("bnji,bnj->bni", RL, s["F"][..., :3]), np.einsum("bnji,bnj->bni", RR, s["F"][..., 3:])], -1)
    assert np.abs(rec["messages"]["F_local"].cpu().numpy() - Floc).max() < 1e-9 * np.abs(Floc).max()
    # torque / Euler / thermal rows of every node against the oracle's reference-mode node
    W = np.concatenate([s["F"][..., :3], np.zeros((B, N, 3)), s["F"][..., 3:], np.zeros((B, N, 3))], -1).reshape(B * N, 12).T
    tau, qnext, Tnext = orc.node_eval_ref([fl, fr], +1.0, qn, np.ascontiguousarray(s["qd"].reshape(B * N, nq).T),
                                          np.ascontiguousarray(W), np.ascontiguousarray(s["T"][:, :N].reshape(B * N, nq).T), drv.h)
    got = rec["rows"]
    assert np.abs(got["tau"].cpu().numpy() - tau.T.reshape(B, N, nq)).max() < 1e-9 * np.abs(tau).max()
    assert np.abs(got["T_defect"].cpu().numpy() - (Tnext.T.reshape(B, N, nq) - s["T"][:, 1:])).max() < 1e-9 * np.abs(Tnext).max()
    # rows that couple the two arms
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pL, pR = pL.T.reshape(B, N, 3), pR.T.reshape(B, N, 3)
    rel = np.einsum("bnji,bnj->bni", RL, pR - pL)
    rp0, ro0, box = rng.normal(size=(B, 3)), rng.normal(size=(B, 3)) * 0.1, np.array([0.6, 0.0, 1.0])
    rows = drv.nodes.box_rows(dev(s["q"]), dev(s["qd"]), dev(s["F"][..., :3]), dev(s["F"][..., 3:]), 5.0, rp0, ro0, box)
    ref_rel = rel - np.concatenate([rp0[:, None], rel[:, :-1]], 1)
    Ro = RL @ RR.transpose(0, 1, 3, 2)
    sk = 0.5 * (Ro - Ro.transpose(0, 1, 3, 2))
    ref_e = np.stack([sk[..., 2, 1], sk[..., 2, 0], sk[..., 1, 0]], -1) - ro0[:, None]
    FL, FR = s["F"][..., :3], s["F"][..., 3:]
    ref_m = np.cross(pL - pR, FL) + np.cross(pR - pL, FR)
    ref_cost = (1000 * ((0.5 * (pL + pR) - box) ** 2).sum(-1) + 100 * (s["qd"] ** 2).sum(-1) + 10 * (FL ** 2).sum(-1) + 10 * (FR ** 2).sum(-1)).sum(-1)
    for key, ref in (("rel_pos", ref_rel), ("rel_ori", ref_e), ("moment_eq", ref_m), ("cost", ref_cost),
                     ("force_eq", FL + FR - np.array([0, 0, 9.81 * 5.0]))):
        assert np.abs(rows[key].cpu().numpy() - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), key
