"""Config C3's coupled fatigue (two arms sharing one box load; include/mpcf.h: mpcf_model_set_coupling, csrc/kernels_couple.cu,
oracle/core.inc.h: step_rk4_coupled).  Builder-defined, parity unpinned by the reference; pinned here by
  CPU: weight = 0 reduces to the uncoupled step; complex-step Jacobian = central differences; the cross-arm block is there;
  GPU: states and every Jacobian plane against the oracle at 1e-9 (closed-form blocks vs complex-step through the whole step)."""
import numpy as np
import pytest

from conftest import oracle_model_from_export, random_inputs, rel_err_rows
from mpc_fatigue_b200.coupling import box_load_coupling
from mpc_fatigue_b200.model import Model, data_urdf
from oracle.pyoracle import Oracle

TOL = 1e-9


def _setup(kind):
    if kind == "pilz6x2":
        m = Model.from_urdf(data_urdf("pilz6x2"), armature=1e-2)
    else:
        m = Model.synthetic("dual_arm", 14, seed=3, armature=1e-2)
    cpl = box_load_coupling(m)
    return m, cpl, Oracle(oracle_model_from_export(m))


def test_oracle_coupled_step_consistency():
    m, (f0, f1, w), orc = _setup("pilz6x2")
    om = orc.model
    U, dt, n = 5, 0.02, 12
    q, qd, tau, f, _ = random_inputs(om, U, seed=8)
    # no box: the uncoupled step
    a = orc.step_rk4_coupled((f0, f1), 0.0, q, qd, tau, f, dt)
    b = orc.step_rk4(q, qd, tau, f, dt)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    qn, qdn, fn, J = orc.step_rk4_coupled((f0, f1), w, q, qd, tau, f, dt, jac=True)
    assert np.abs(qn - b[0]).max() < 1e-13 and np.abs(qdn - b[1]).max() < 1e-12 and np.abs(fn - b[2]).max() > 1e-6  # dynamics untouched, fatigue not
    # central differences of the coupled step along every state / control direction
    x0 = [q, qd, tau, f]
    for blk in range(4):
        for j in range(n):
            h = 1e-6 * max(1.0, float(np.abs(x0[blk][j]).max()))
            xp, xm = [a.copy() for a in x0], [a.copy() for a in x0]
            xp[blk][j] += h
            xm[blk][j] -= h
            fp = np.concatenate(orc.step_rk4_coupled((f0, f1), w, *xp, dt))
            fm = np.concatenate(orc.step_rk4_coupled((f0, f1), w, *xm, dt))
            fd = (fp - fm) / (2 * h)
            col = J[:, blk * n + j, :]
            assert np.abs(col - fd).max() < 2e-6 * max(1.0, np.abs(fd).max()), (blk, j)
    # the cross-arm block d f+_arm0 / d f_arm1 is structurally non-zero, d (q+, qd+) / d f stays zero
    assert np.abs(J[2 * n + 1:2 * n + 4, 3 * n + 6:4 * n]).min() > 1e-9  # (joint 1's axis is vertical: the box weight has no moment about it)
    assert not J[:2 * n, 3 * n:4 * n].any()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["pilz6x2", "dual_arm14"])
def test_gpu_coupled_step_and_jacobian(kind):
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    m, cpl, orc = _setup(kind)
    om, n = orc.model, m.n
    U, dt = 161, 0.02
    q, qd, tau, f, _ = random_inputs(om, U, seed=9)
    d = [torch.from_numpy(a).cuda() for a in (q, qd, tau, f)]
    ev = BatchEvaluator(m)
    base = [t.cpu().numpy() for t in ev.step_rk4_jvp(*d, dt)]  # before the coupling is set: the decoupled result
    m.set_coupling(cpl)
    rq, rqd, rf, rj = orc.step_rk4_coupled(cpl[:2], cpl[2], q, qd, tau, f, dt, jac=True)
    gq, gqd, gf = [t.cpu().numpy() for t in ev.step_rk4(*d, dt)]
    assert rel_err_rows(gq, rq) < TOL and rel_err_rows(gqd, rqd) < TOL and rel_err_rows(gf, rf) < TOL
    got = [t.cpu().numpy() for t in ev.step_rk4_jvp(*d, dt)]
    assert rel_err_rows(got[0], rq) < TOL and rel_err_rows(got[1], rqd) < TOL and rel_err_rows(got[2], rf) < TOL
    P = 4 * n + 1
    assert rel_err_rows(got[3].reshape(3 * n * P, U), rj.reshape(3 * n * P, U)) < TOL
    # the dynamics rows are those of the decoupled model, bit for bit; the fatigue rows changed
    assert np.array_equal(got[3][:2 * n], base[3][:2 * n]) and np.array_equal(got[0], base[0]) and np.array_equal(got[1], base[1])
    assert np.abs(got[3][2 * n:, 3 * n:4 * n] - base[3][2 * n:, 3 * n:4 * n]).max() > 1e-8
    # per-unit dt and a ragged batch through the same path
    dt_u = np.ascontiguousarray(np.random.default_rng(5).uniform(0.005, 0.03, U))
    rq, rqd, rf, rj = orc.step_rk4_coupled(cpl[:2], cpl[2], q, qd, tau, f, 0.0, dt_u=dt_u, jac=True)
    got = [t.cpu().numpy() for t in ev.step_rk4_jvp(*d, torch.from_numpy(dt_u).cuda())]
    assert rel_err_rows(got[2], rf) < TOL and rel_err_rows(got[3].reshape(3 * n * P, U), rj.reshape(3 * n * P, U)) < TOL
    # entries that cannot honour the coupling refuse instead of ignoring it
    with pytest.raises(ValueError):
        ev.step_rk4_jvp(*d, dt, direct=True)
    m.set_coupling(None)
    again = ev.step_rk4_jvp(*d, dt)[3].cpu().numpy()
    assert np.array_equal(again, base[3])
