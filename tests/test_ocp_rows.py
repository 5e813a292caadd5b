"""Fused reference-mode OCP node rows (include/mpcf.h: mpcf_ocp_rows_batch, csrc/kernels_rows.cu; oracle/core.inc.h: ocp_rows).

CPU: the oracle's row function is pinned on the reference's stored dual-arm solutions (plotter/Result_* via
tests/golden/plotter_solutions.npz): on the single-tree URDF of both arms the force / moment / distance / Euler rows vanish to
IPOPT's tolerance and the torque rows sit on the active bounds; the two-arm rows also agree with an independent numpy
composition of the bridge Functions (the batched-torch code they replace in ocp.py), and the complex-step kinematic Jacobian
with central differences.  GPU: rows, cost and derivative outputs against the oracle for one- and two-arm families."""
import os

import numpy as np
import pytest

from conftest import ROOT, oracle_model_from_export, random_inputs, rel_err_rows
from mpc_fatigue_b200.model import Model, data_urdf
from mpc_fatigue_b200.ocp import DualArmBoxOCP
from oracle.pyoracle import Oracle
from oracle.urdf_model import load_urdf

GOLD = os.path.join(ROOT, "tests", "golden")
_T = lambda a: np.ascontiguousarray(np.asarray(a).T)


def _two_arm_inputs(om, B, N, seed):
    U = B * N
    q, qd, _, T, _ = random_inputs(om, U, seed=seed)
    rng = np.random.default_rng(seed + 1)
    F = np.ascontiguousarray(rng.uniform(-60, 60, (6, U)))
    return q, qd, F, T, np.ascontiguousarray(rng.uniform(-1, 1, (om.n, B))), np.ascontiguousarray(rng.uniform(20, 80, (om.n, B))), \
        np.ascontiguousarray(rng.uniform(-0.5, 0.5, (3, B))), np.ascontiguousarray(rng.uniform(-0.5, 0.5, (3, B)))


OPTS2 = dict(narm=2, wsign=+1.0, fdes=(0.0, 0.0, 9.81 * 30), dist2_ref=0.04, mu=0.5, p_ref=(0.3, 0.6, 0.4), w_box=1000.0, w_qd=100.0, w_F=10.0, h=0.5)


def test_oracle_rows_on_the_reference_solutions():
    """Box_Pilz_6DOF2.py solutions replayed through ocp_rows on urdf/2_pilz_robot_6DOF.urdf (both arms in one tree)."""
    sols = np.load(os.path.join(GOLD, "plotter_solutions.npz"))
    om = load_urdf(data_urdf("pilz6x2"))
    orc = Oracle(om)
    ee = [om.frame_names.index("end_effector"), om.frame_names.index("sec_end_effector")]
    for key, tol_def in (("Result_1", 1e-12), ("Result_2", 1e-12), ("Result_4", 1e-9), ("plotter", 1e-9)):
        s = DualArmBoxOCP.parse_solution(sols[key])
        N = s["N"]
        h = 2.0 / N
        opts = dict(narm=2, ee_frame=ee, wsign=-1.0, fdes=(0.0, 0.0, 9.81 * 30), dist2_ref=0.04, h=h)
        F = np.ascontiguousarray(np.hstack([s["F_LR"], s["F_RR"]]).T)
        rows, cost = orc.ocp_rows(opts, 1, N, _T(s["q"][:N]), _T(s["qd"]), F, q_last=_T(s["q"][N:N + 1]))
        assert np.abs(rows[0:3]).max() < 1.1e-4 and np.abs(rows[3:6]).max() < 1.1e-4, key  # equilibrium rows within +-pos_toll = 1e-4
        assert np.abs(rows[6]).max() < 1e-7, key                                              # |E1 - E2|^2 = 0.04
        assert np.abs(rows[26 + 12:26 + 24]).max() < tol_def, key                            # Euler defects
        tau = rows[26:26 + 12]
        if key == "Result_2":   # right arm on the final-third bounds (-5, 5, 5): Box_Pilz_6DOF.py
            assert np.abs(tau[6:9, 53:] - np.array([[-5.0], [5.0], [5.0]])).max() < 4e-7
        if key == "Result_4":   # left arm: tau_LR0 = tau_LR1 = +5 active
            assert np.abs(tau[0:2, 53:] - 5.0).max() < 2e-7


def test_oracle_two_arm_rows_match_a_numpy_composition_of_the_bridge_functions():
    om = load_urdf(data_urdf("pilz6x2"), armature=1e-2)
    orc = Oracle(om)
    ee = [om.frame_names.index("end_effector"), om.frame_names.index("sec_end_effector")]
    B, N, n = 3, 4, 12
    q, qd, F, T, q_last, T_last, rp0, ro0 = _two_arm_inputs(om, B, N, 50)
    opts = dict(OPTS2, ee_frame=ee)
    rows, cost = orc.ocp_rows(opts, B, N, q, qd, F, T, q_last, T_last, rp0, ro0)
    U = B * N
    pL, RL = orc.fk(ee[0], q)
    pR, RR = orc.fk(ee[1], q)
    RL, RR = RL.reshape(3, 3, U), RR.reshape(3, 3, U)
    FL, FR = F[:3], F[3:]
    assert np.allclose(rows[0:3], FL + FR - np.array(opts["fdes"])[:, None], atol=1e-12)
    assert np.allclose(rows[3:6], np.cross(pL - pR, FL, axis=0) + np.cross(pR - pL, FR, axis=0), atol=1e-10)
    rel = np.einsum("jiu,ju->iu", RL, pR - pL)
    prev = np.concatenate([rp0, rel[:, :U - B]], axis=1)
    assert np.allclose(rows[7:10], rel - prev, atol=1e-12)
    Ro = np.einsum("iku,jku->iju", RL, RR)
    sk = 0.5 * (Ro - Ro.transpose(1, 0, 2))
    e = np.stack([sk[2, 1], sk[2, 0], sk[1, 0]])
    assert np.allclose(rows[10:13], e - np.tile(ro0, (1, N)), atol=1e-12)
    f1, f2 = -np.einsum("jiu,ju->iu", RL, FL), -np.einsum("jiu,ju->iu", RR, FR)
    mu = opts["mu"]
    A1 = np.array([[0, -1, 0], [0, -mu, 1], [0, -mu, -1], [1, -mu, 0], [-1, -mu, 0]])   # confriction.py:251-256
    A2 = np.array([[0, 1, 0], [1, mu, 0], [-1, mu, 0], [0, mu, 1], [0, mu, -1]])        # :262-266
    assert np.allclose(rows[13:18], A1 @ f1, atol=1e-10) and np.allclose(rows[18:23], A2 @ f2, atol=1e-10)
    pbox = 0.5 * (pL + pR)
    W = np.vstack([FL, np.zeros((3, U)), FR, np.zeros((3, U))])
    tau, qn, Tn = orc.node_eval_ref(ee, +1.0, q, qd, np.ascontiguousarray(W), T, opts["h"])
    assert np.allclose(rows[26:38], tau, atol=1e-10)
    qnext = np.concatenate([q[:, B:], q_last], axis=1)
    Tnext = np.concatenate([T[:, B:], T_last], axis=1)
    assert np.allclose(rows[38:50], qn - qnext, atol=1e-12) and np.allclose(rows[50:62], Tn - Tnext, atol=1e-10)
    want = 1000 * ((pbox - np.array(opts["p_ref"])[:, None]) ** 2).sum(0) + 100 * (qd ** 2).sum(0) + 10 * (F ** 2).sum(0)  # mpc_principal.py:281-284
    assert np.allclose(cost, want, rtol=1e-13)


def test_oracle_kinematic_jacobian_against_central_differences():
    om = load_urdf(data_urdf("pilz6x2"))
    orc = Oracle(om)
    ee = [om.frame_names.index("end_effector"), om.frame_names.index("sec_end_effector")]
    B, N, n = 2, 2, 12
    q, qd, F, T, q_last, T_last, rp0, ro0 = _two_arm_inputs(om, B, N, 51)
    opts = dict(OPTS2, ee_frame=ee)
    _, _, J = orc.ocp_rows(opts, B, N, q, qd, F, kin_jac=True)
    for d in range(n + 6):
        xp, xm = [q.copy(), F.copy()], [q.copy(), F.copy()]
        blk, i = (0, d) if d < n else (1, d - n)
        eps = 1e-6 * max(1.0, np.abs(xp[blk][i]).max())
        xp[blk][i] += eps
        xm[blk][i] -= eps
        # the relative-position row also moves with the previous node's q: perturb one node at a time
        for k in range(N):
            sel = slice(k * B, (k + 1) * B)
            qp, qm, Fp, Fm = q.copy(), q.copy(), F.copy(), F.copy()
            qp[:, sel], qm[:, sel], Fp[:, sel], Fm[:, sel] = xp[0][:, sel], xm[0][:, sel], xp[1][:, sel], xm[1][:, sel]
            rp, _ = orc.ocp_rows(opts, B, N, qp, qd, Fp)
            rm, _ = orc.ocp_rows(opts, B, N, qm, qd, Fm)
            fd = (rp[:26, sel] - rm[:26, sel]) / (2 * eps)
            assert np.abs(fd - J[:, d, sel]).max() < 5e-6 * max(1.0, np.abs(fd).max()), (d, k)


GPU_CASES = [("pilz6", 1), ("pilz3", 1), ("pilz6x2", 2), ("dual_arm14", 2)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,narm", GPU_CASES)
def test_gpu_ocp_rows_against_the_oracle(name, narm):
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    m = Model.synthetic("dual_arm", 14, seed=3, armature=1e-2) if name == "dual_arm14" else Model.from_urdf(data_urdf(name), armature=1e-2)
    om = oracle_model_from_export(m)
    orc, ev = Oracle(om), BatchEvaluator(m)
    n, B, N = m.n, 37, 5
    U = B * N
    fpar = m.export("fparent").tolist()
    last = lambda j: max(i for i, p in enumerate(fpar) if p == j)
    ee = [last(n - 1)] if narm == 1 else [last(n // 2 - 1), last(n - 1)]
    q, qd, F6, T, q_last, T_last, rp0, ro0 = _two_arm_inputs(om, B, N, 60)
    F = np.ascontiguousarray(F6[:3 * narm])
    opts = dict(OPTS2, narm=narm, ee_frame=ee, wsign=-1.0 if narm == 1 else +1.0, w_F=-1.0 if narm == 1 else 10.0)
    rrows, rcost, rJ = orc.ocp_rows(opts, B, N, q, qd, F, T, q_last, T_last, rp0 if narm == 2 else None, ro0 if narm == 2 else None, kin_jac=True)
    g = lambda a: None if a is None else torch.from_numpy(a).cuda()
    kw = {k: opts[k] for k in ("wsign", "fdes", "dist2_ref", "mu", "p_ref", "w_box", "w_qd", "w_F", "h")}
    rows, cost, dF, dT, kj = ev.ocp_rows(B, N, ee, g(q), g(qd), g(F), g(T), g(q_last), g(T_last), g(rp0) if narm == 2 else None,
                                         g(ro0) if narm == 2 else None, derivatives=True, **kw)
    assert rows.shape[0] == (26 if narm == 2 else 3) + 3 * n
    assert rel_err_rows(rows.cpu().numpy(), rrows) < 1e-9
    assert np.abs(cost.cpu().numpy() - rcost).max() < 1e-9 * np.abs(rcost).max()
    kin = rows.shape[0] - 3 * n
    assert rel_err_rows(kj.cpu().numpy().reshape(-1, U), rJ.reshape(-1, U)) < 1e-9
    # d tau / d F = wsign J_lin^T of every arm's end-effector (frame Jacobian rows 0-2), zero across arms
    dFh = dF.cpu().numpy()
    for c in range(narm):
        J = orc.jacobian(ee[c], q).reshape(6, n, U)
        assert np.abs(dFh[:, 3 * c:3 * c + 3] - opts["wsign"] * J[:3].transpose(1, 0, 2)).max() < 1e-11
    # d T+ / d tau by central differences of the oracle's ZOH map at the computed torque
    tau = rrows[kin:kin + n]
    eps = 1e-4
    fd = (orc.fatigue_zoh(T, tau + eps, qd, opts["h"]) - orc.fatigue_zoh(T, tau - eps, qd, opts["h"])) / (2 * eps)
    assert np.abs(dT.cpu().numpy() - fd).max() < 1e-7 * max(1.0, np.abs(fd).max())
    # optional inputs absent: thermal rows and last-node defects are zero, nothing else changes
    rows2, cost2 = ev.ocp_rows(B, N, ee, g(q), g(qd), g(F), **kw)
    r2 = rows2.cpu().numpy()
    assert not r2[kin + 2 * n:].any() and not r2[kin + n:kin + 2 * n, (N - 1) * B:].any()
    assert np.array_equal(r2[kin:kin + n], rows.cpu().numpy()[kin:kin + n])


@pytest.mark.gpu
def test_gpu_fused_box_ocp_agrees_with_the_per_function_composition():
    """FusedBoxOCP (one launch) against ThermalMPCNodes.evaluate + box_rows (node_eval_ref / fk launches composed in torch)."""
    import torch
    from mpc_fatigue_b200.ocp import FusedBoxOCP, ThermalMPCNodes
    m = Model.synthetic("dual_arm", 14, seed=3, armature=1e-2)
    n, B, N = 14, 9, 6
    fpar = m.export("fparent").tolist()
    last = lambda j: max(i for i, p in enumerate(fpar) if p == j)
    names = [m.frame_names[last(6)], m.frame_names[last(13)]]
    g = torch.Generator(device="cuda").manual_seed(5)
    r = lambda *s: torch.rand(*s, dtype=torch.float64, device="cuda", generator=g)
    q, qd = 2 * r(B, N + 1, n) - 1, r(B, N, n) - 0.5
    FL, FR = 40 * (r(B, N, 3) - 0.5), 40 * (r(B, N, 3) - 0.5)
    Temp = 20 + 40 * r(B, N + 1, n)
    rp0, ro0, box0 = r(B, 3) - 0.5, r(B, 3) - 0.5, (0.2, 0.1, 0.5)
    fused = FusedBoxOCP(m, names, p_ref=box0, T=20.0, N=N).evaluate(q, qd, FL, FR, Temp, rp0, ro0)
    old = ThermalMPCNodes(m, names, T=20.0, N=N)
    W = torch.cat([FL, torch.zeros_like(FL), FR, torch.zeros_like(FR)], dim=-1)
    a = old.evaluate(q, qd, W, Temp)
    b = old.box_rows(q, qd, FL, FR, 30.0, rp0, ro0, box0)
    close = lambda x, y: float((x - y).abs().max()) < 1e-9 * max(1.0, float(y.abs().max()))
    assert close(fused["tau"], a["tau"]) and close(fused["q_defect"], a["q_defect"]) and close(fused["T_defect"], a["T_defect"])
    assert close(fused["force_eq"], b["force_eq"]) and close(fused["moment_eq"], b["moment_eq"])
    assert close(fused["rel_pos"], b["rel_pos"]) and close(fused["rel_ori"], b["rel_ori"]) and close(fused["cost"], b["cost"])
