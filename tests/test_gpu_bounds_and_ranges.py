"""GPU: torque-bound residuals from a per-node table (F0 decaying, F2 stepwise, F3 switch-off) and the unit-range form of the
Jacobian entry (one chunk-sized Jacobian buffer reused over a larger batch)."""
import numpy as np
import pytest

from mpc_fatigue_b200.model import Model, data_urdf
from mpc_fatigue_b200.ocp import (bound_table_from, f0_bound_schedule, f0_bound_table, step_bound_table, switch_off_bound_table)


def test_switch_off_table_follows_the_reference_rule():
    # python/Centauro_script/Centauro_dynamics.py:327-348 with S / C of CentaurOCP.py:64-71, written out literally
    N, n = 9, 4
    lbt, ubt = [-147.0, -147.0, -55.0, -28.32], [147.0, 147.0, 55.0, 28.32]
    S, C = [1, 0, 1, 0], [0.0, 0.0, 3.0, 0.0]
    tb = switch_off_bound_table(N, lbt, ubt, S, C)
    for k in range(N):
        for i in range(n):
            if S[i] == 1 and not (k < int(N / 3)):
                want = (-C[i], C[i])
            else:
                want = (lbt[i], ubt[i])
            assert tuple(tb[k, i]) == want, (k, i)


def test_f0_table_matches_the_schedule():
    tb = f0_bound_table(60, 6, 2.0 / 60)
    b = f0_bound_schedule(60, 2.0 / 60)
    assert tb.shape == (60, 6, 2) and np.array_equal(tb[:, 3, 1], b) and np.array_equal(tb[:, 0, 0], -b)


@pytest.mark.gpu
@pytest.mark.parametrize("mech", ["F0", "F2", "F3"])
def test_gpu_bound_table_residual(mech):
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.synth import synth_batch
    m = Model.from_urdf(data_urdf("pilz6"), armature=1e-2)
    ev = BatchEvaluator(m)
    n, B, N, dt = 6, 517, 30, 2.0 / 30
    lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
    q, qd, tau, f = synth_batch(lim, 0, B, N, device="cuda")
    qn, qdn, fn = ev.step_rk4(q, qd, tau, f, dt)
    if mech == "F0":
        tb = f0_bound_table(N, n, dt, 50.0, 2.0, 15.0)
    elif mech == "F2":  # thirds of the horizon, asymmetric bounds (Box_Pilz_6DOF.py final-third bounds: [-10, 5])
        lb, ub = step_bound_table(N, [(np.full(n, -50.0), np.full(n, 50.0)), (np.full(n, -20.0), np.full(n, 30.0)), (np.full(n, -10.0), np.full(n, 5.0))])
        tb = bound_table_from(lb, ub)
    else:
        tb = switch_off_bound_table(N, -m.export("tau_max"), m.export("tau_max"), [1, 1, 0, 0, 1, 0], [0.0, 2.0, 0.0, 0.0, 1.0, 0.0])
    red = ev.cost_residual_table(B, N, q, qd, f, tau, qn, qdn, fn, torch.from_numpy(tb).cuda(), w_qd=1.0, w_tau=1e-2, f_max=60.0)
    # numpy restatement of the four rows
    t = tau.cpu().numpy().reshape(n, N, B)
    v = qd.cpu().numpy().reshape(n, N, B)
    viol = np.maximum(np.maximum(tb[:, :, 0].T[:, :, None] - t, t - tb[:, :, 1].T[:, :, None]), 0.0).max(axis=(0, 1))
    cost = (v * v + 1e-2 * t * t).sum(axis=(0, 1))
    r = red.cpu().numpy()
    assert np.abs(r[2] - viol).max() < 1e-12 and np.abs(r[0] - cost).max() < 1e-9 * np.abs(cost).max()
    fnh = fn.cpu().numpy().reshape(n, N, B)
    assert np.abs(r[3] - np.maximum(fnh - 60.0, 0.0).max(axis=(0, 1))).max() < 1e-12
    if mech == "F0":  # the table entry reproduces the closed-form kernel exactly
        old = ev.cost_residual(B, N, q, qd, f, tau, qn, qdn, fn, dt, f_max=60.0)
        assert torch.equal(old[1], red[1]) and torch.equal(old[3], red[3]) and float((old[2] - red[2]).abs().max()) < 1e-12
    # chunk form: two scenario halves write their columns of one [4, B] send buffer... (units are node-major per chunk batch)
    out = torch.zeros((4, B + 3), dtype=torch.float64, device="cuda")
    ev.cost_residual_table(B, N, q, qd, f, tau, qn, qdn, fn, torch.from_numpy(tb).cuda(), f_max=60.0, out=out, out_col0=3)
    assert torch.equal(out[:, 3:], red) and not bool(out[:, :3].any())


@pytest.mark.gpu
def test_gpu_unit_range_entry_reuses_one_jacobian_buffer():
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.synth import synth_batch
    for m in (Model.from_urdf(data_urdf("pilz6"), armature=1e-2), Model.synthetic("chain", 5, seed=2, armature=1e-2)):
        ev = BatchEvaluator(m)
        n, U = m.n, 333
        lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
        q, qd, tau, f = synth_batch(lim, 0, U, 1, device="cuda")
        ref = ev.step_rk4_jvp(q, qd, tau, f, 0.02)
        out = tuple(torch.zeros_like(q) for _ in range(3))
        jac = torch.empty((3 * n, 4 * n + 1, 128), dtype=torch.float64, device="cuda")
        for u0 in range(0, U, 100):
            cnt = min(100, U - u0)
            ev.step_rk4_jvp_range(q, qd, tau, f, 0.02, u0, cnt, out, jac)
            assert torch.equal(jac[:, :, :cnt], ref[3][:, :, u0:u0 + cnt]), (m.kernel_family, u0)
        assert all(torch.equal(a, b) for a, b in zip(out, ref[:3]))
