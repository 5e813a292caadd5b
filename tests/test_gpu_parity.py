"""GPU-vs-oracle parity through the C-ABI (run with -m gpu on a B200).

Tolerance: the north star asks for <= 1e-9 relative on states, torques and Jacobians (fp64).
"""
import numpy as np
import pytest

from conftest import oracle_model_from_export, random_inputs, rel_err, rel_err_rows

pytestmark = pytest.mark.gpu
TOL = 1e-9

CASES = [
    ("pilz6", dict(urdf="pilz6", armature=1e-2)),
    ("pilz3", dict(urdf="pilz3", armature=0.0)),
    ("pilz6x2", dict(urdf="pilz6x2", armature=1e-2)),
    ("pilz6_second", dict(urdf="pilz6_second", armature=1e-2)),
    ("chain7", dict(synthetic=("chain", 7, 2), armature=1e-3)),
    ("dualarm14", dict(synthetic=("dual_arm", 14, 4), armature=1e-2)),
    ("chain10", dict(synthetic=("chain", 10, 3), armature=1e-3)),
    ("humanoid37", dict(synthetic=("humanoid", 37, 7), armature=1e-2)),
]


@pytest.fixture(scope="module", params=CASES, ids=[c[0] for c in CASES])
def setup(request):
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import Model, data_urdf
    from oracle.pyoracle import Oracle
    name, spec = request.param
    if "urdf" in spec:
        m = Model.from_urdf(data_urdf(spec["urdf"]), armature=spec["armature"])
    else:
        kind, ndof, seed = spec["synthetic"]
        m = Model.synthetic(kind, ndof, seed=seed, armature=spec["armature"])
    om = oracle_model_from_export(m)
    U = 257 if m.n <= 12 else 130  # ragged against the 128-thread block
    arrs = random_inputs(om, U, seed=11)
    dev = [torch.from_numpy(a).cuda() for a in arrs]
    return dict(name=name, m=m, om=om, orc=Oracle(om), ev=BatchEvaluator(m), U=U, host=arrs, dev=dev, torch=torch)


def test_rnea(setup):
    q, qd, tau, f, qdd = setup["host"]
    dq, dqd, dtau, df, dqdd = setup["dev"]
    ref = setup["orc"].rnea(q, qd, qdd)
    got = setup["ev"].rnea(dq, dqd, dqdd).cpu().numpy()
    assert rel_err_rows(got, ref) < TOL
    ref0 = setup["orc"].rnea(q, qd, None)
    got0 = setup["ev"].rnea(dq, dqd, None).cpu().numpy()
    assert rel_err_rows(got0, ref0) < TOL


def test_fk_and_jacobian_every_frame(setup):
    q = setup["host"][0]
    dq = setup["dev"][0]
    m = setup["m"]
    frames = range(m.nframes) if m.n <= 12 else list(range(0, m.nframes, 7)) + [m.nframes - 1]
    for fr in frames:
        pos, rot = setup["orc"].fk(fr, q)
        gpos, grot = setup["ev"].fk(fr, dq)
        assert rel_err_rows(gpos.cpu().numpy(), pos) < TOL and rel_err_rows(grot.cpu().numpy(), rot) < TOL
        J = setup["orc"].jacobian(fr, q)
        gJ = setup["ev"].jacobian(fr, dq).cpu().numpy()
        assert rel_err_rows(gJ, J) < TOL


def test_jac_t_wrench_and_node_eval(setup):
    torch = setup["torch"]
    q, qd, tau, f, qdd = setup["host"]
    dq, dqd, dtau, df, dqdd = setup["dev"]
    m, U = setup["m"], setup["U"]
    rng = np.random.default_rng(5)
    ee = [m.nframes - 1] if m.n != 12 else [m.frame_id("end_effector"), m.frame_id("sec_end_effector")]
    W = np.ascontiguousarray(rng.uniform(-50, 50, (6 * len(ee), U)))
    dW = torch.from_numpy(W).cuda()
    h = 0.5
    for wsign in (-1.0, 1.0):
        rt, rq, rT = setup["orc"].node_eval_ref(ee, wsign, q, qd, W, f, h, qdd=qdd)
        gt, gq, gT = setup["ev"].node_eval_ref(ee, wsign, dq, dqd, dW, df, h, qdd=dqdd)
        assert rel_err_rows(gt.cpu().numpy(), rt) < TOL
        assert rel_err_rows(gq.cpu().numpy(), rq) < TOL
        assert rel_err_rows(gT.cpu().numpy(), rT) < TOL
    # J^T W alone == J^T W from the materialised Jacobian of the oracle
    J = setup["orc"].jacobian(ee[0], q).reshape(6, m.n, U)
    ref = np.einsum("rnu,ru->nu", J, W[:6])
    got = setup["ev"].jac_t_wrench(ee[0], dq, dW[:6].contiguous()).cpu().numpy()
    assert rel_err_rows(got, ref) < TOL


def test_node_eval_ref_jvp(setup):
    """Jacobian blocks of the reference-mode torque rows vs complex-step through the oracle's node evaluation."""
    torch = setup["torch"]
    q, qd, tau, f, qdd = setup["host"]
    dq, dqd, dtau, df, dqdd = setup["dev"]
    m = setup["m"]
    U = min(setup["U"], 40)
    rng = np.random.default_rng(6)
    ee = [m.nframes - 1] if m.n != 12 else [m.frame_id("end_effector"), m.frame_id("sec_end_effector")]
    W = np.ascontiguousarray(rng.uniform(-50, 50, (6 * len(ee), U)))
    sl = lambda a: np.ascontiguousarray(a[:, :U])
    dsl = lambda a: a[:, :U].contiguous()
    for wsign in (-1.0, 1.0):
        rDq, rDv = setup["orc"].node_eval_ref_jvp(ee, wsign, sl(q), sl(qd), W, qdd=sl(qdd))
        gDq, gDv = setup["ev"].node_eval_ref_jvp(ee, wsign, dsl(dq), dsl(dqd), torch.from_numpy(W).cuda(), qdd=dsl(dqdd))
        assert rel_err_rows(gDq.cpu().numpy(), rDq) < TOL
        assert rel_err_rows(gDv.cpu().numpy(), rDv) < TOL
    # without wrenches it reduces to the inverse-dynamics derivatives
    gDq0, gDv0 = setup["ev"].node_eval_ref_jvp([], 1.0, dsl(dq), dsl(dqd), None, qdd=dsl(dqdd))
    aDq, aDv, _ = setup["ev"].rnea_derivs(dsl(dq), dsl(dqd), dsl(dqdd))
    assert rel_err(gDq0.cpu().numpy(), aDq.cpu().numpy()) < 1e-10 and rel_err(gDv0.cpu().numpy(), aDv.cpu().numpy()) < 1e-10


def test_aba(setup):
    q, qd, tau, f, qdd = setup["host"]
    dq, dqd, dtau, df, dqdd = setup["dev"]
    ref = setup["orc"].aba(q, qd, tau)
    got = setup["ev"].aba(dq, dqd, dtau).cpu().numpy()
    assert rel_err_rows(got, ref) < TOL
    # round trip through the GPU RNEA: RNEA(q, qd, ABA(q, qd, tau)) == tau
    back = setup["ev"].rnea(dq, dqd, setup["ev"].aba(dq, dqd, dtau)).cpu().numpy()
    assert rel_err_rows(back, tau) < 1e-8


def test_step_rk4(setup):
    torch = setup["torch"]
    q, qd, tau, f, qdd = setup["host"]
    dq, dqd, dtau, df, dqdd = setup["dev"]
    for dt in (0.02, 0.005):
        rq, rqd, rf = setup["orc"].step_rk4(q, qd, tau, f, dt)
        gq, gqd, gf = setup["ev"].step_rk4(dq, dqd, dtau, df, dt)
        assert rel_err_rows(gq.cpu().numpy(), rq) < TOL
        assert rel_err_rows(gqd.cpu().numpy(), rqd) < TOL
        assert rel_err_rows(gf.cpu().numpy(), rf) < TOL
    # per-unit dt
    dtu = np.ascontiguousarray(np.random.default_rng(2).uniform(0.001, 0.03, setup["U"]))
    rq, rqd, rf = setup["orc"].step_rk4(q, qd, tau, f, 0.0, dt_u=dtu)
    gq, gqd, gf = setup["ev"].step_rk4(dq, dqd, dtau, df, torch.from_numpy(dtu).cuda())
    assert rel_err_rows(gqd.cpu().numpy(), rqd) < TOL and rel_err_rows(gf.cpu().numpy(), rf) < TOL


def test_rollout_rk4(setup):
    """Single-shooting rollout == N chained oracle steps; and == chaining the GPU's own per-node step kernel bit for bit."""
    torch = setup["torch"]
    q, qd, tau, f, qdd = setup["host"]
    m = setup["m"]
    n, B, N, dt = m.n, 9, 5, 0.004
    rng = np.random.default_rng(21)
    q0, qd0, f0 = (np.ascontiguousarray(a[:, :B]) for a in (q, 0.3 * qd, f))
    taus = np.ascontiguousarray(np.tile(tau[:, :B], (1, N)) * rng.uniform(0.5, 1.0, (n, N * B)))
    d = [torch.from_numpy(a).cuda() for a in (q0, qd0, f0, taus)]
    gq, gqd, gf = setup["ev"].rollout_rk4(d[0], d[1], d[2], d[3], N, dt)
    x = (q0, qd0, f0)
    xs = (d[0], d[1], d[2])
    for k in range(N):
        tk = np.ascontiguousarray(taus[:, k * B:(k + 1) * B])
        x = setup["orc"].step_rk4(x[0], x[1], tk, x[2], dt)
        xs = setup["ev"].step_rk4(xs[0], xs[1], d[3][:, k * B:(k + 1) * B].contiguous(), xs[2], dt)
        for got, ref, own in zip((gq, gqd, gf), x, xs):
            blk = got[:, k * B:(k + 1) * B]
            assert rel_err_rows(blk.cpu().numpy(), ref) < 1e-8  # N steps of error growth
            assert torch.equal(blk, own)


def test_fd_derivs(setup):
    q, qd, tau, f, qdd = setup["host"]
    dq, dqd, dtau, df, dqdd = setup["dev"]
    U = min(setup["U"], 40)
    sl = lambda a: np.ascontiguousarray(a[:, :U])
    dsl = lambda a: a[:, :U].contiguous()
    rA, rB, rC = setup["orc"].fd_derivs(sl(q), sl(qd), sl(tau))
    gA, gB, gC = setup["ev"].fd_derivs(dsl(dq), dsl(dqd), dsl(dtau))
    # per plane; on the 37-joint tree the planes that couple different limbs are 10^4-10^5 times smaller than the block maximum
    # and are differences of O(max) terms through M^-1 in BOTH implementations (dual-number ABA here, complex-step ABA in the
    # oracle), so their own-plane relative agreement is limited by conditioning: floor 1e-4 * scale there
    fl = 1e-4 if setup["m"].n > 16 else 1e-6
    assert rel_err_rows(gC.cpu().numpy(), rC, floor=fl) < TOL
    assert rel_err_rows(gA.cpu().numpy(), rA, floor=fl) < TOL
    assert rel_err_rows(gB.cpu().numpy(), rB, floor=fl) < TOL


def test_rnea_derivs(setup):
    q, qd, tau, f, qdd = setup["host"]
    dq, dqd, dtau, df, dqdd = setup["dev"]
    U = min(setup["U"], 40)
    sl = lambda a: np.ascontiguousarray(a[:, :U])
    dsl = lambda a: a[:, :U].contiguous()
    for with_qdd in (True, False):
        rDq, rDv, rM = setup["orc"].rnea_derivs(sl(q), sl(qd), sl(qdd) if with_qdd else None)
        gDq, gDv, gM = setup["ev"].rnea_derivs(dsl(dq), dsl(dqd), dsl(dqdd) if with_qdd else None)
        assert rel_err_rows(gM.cpu().numpy(), rM) < TOL
        assert rel_err_rows(gDq.cpu().numpy(), rDq) < TOL
        assert rel_err_rows(gDv.cpu().numpy(), rDv) < TOL
    # M is the joint-space inertia the oracle's CRBA computes
    n = setup["m"].n
    assert rel_err(gM.cpu().numpy()[:, 0].reshape(n, n), setup["orc"].crba(q[:, 0].copy())) < TOL


@pytest.mark.parametrize("direct", [False, True], ids=["workspace", "direct"])
def test_step_rk4_jvp(setup, direct):
    q, qd, tau, f, qdd = setup["host"]
    dq, dqd, dtau, df, dqdd = setup["dev"]
    U = min(setup["U"], 64 if setup["m"].n > 12 else 257)
    sl = lambda a: np.ascontiguousarray(a[:, :U])
    dsl = lambda a: a[:, :U].contiguous()
    dt = 0.02
    rq, rqd, rf, rj = setup["orc"].step_rk4_jvp(sl(q), sl(qd), sl(tau), sl(f), dt)
    gq, gqd, gf, gj = setup["ev"].step_rk4_jvp(dsl(dq), dsl(dqd), dsl(dtau), dsl(df), dt, direct=direct)
    assert rel_err_rows(gq.cpu().numpy(), rq) < TOL
    assert rel_err_rows(gqd.cpu().numpy(), rqd) < TOL
    assert rel_err_rows(gf.cpu().numpy(), rf) < TOL
    gj = gj.cpu().numpy()
    assert gj.shape == rj.shape
    # every Jacobian entry plane relative to the scale of its own row block
    n = setup["m"].n
    for r0, r1 in ((0, n), (n, 2 * n), (2 * n, 3 * n)):
        assert rel_err_rows(gj[r0:r1], rj[r0:r1]) < TOL, (r0, rel_err_rows(gj[r0:r1], rj[r0:r1]))
    # structure: d(q+,qd+)/df == 0 exactly, df+/df diagonal
    assert np.all(gj[: 2 * n, 3 * n : 4 * n] == 0.0)
    off = gj[2 * n :, 3 * n : 4 * n] * (1 - np.eye(n))[:, :, None]
    assert np.all(off == 0.0)


@pytest.mark.parametrize("urdf", ["pilz6", "pilz6x2"])
def test_step_rk4_jvp_fast_motion(urdf):
    """The derivative kernel rebuilds each link's velocity / acceleration from its child's by undoing the joint; joint
    speeds 8x the URDF limit, 4x the usual torques and a long step stress the cancellation in that recursion."""
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import Model, data_urdf
    from oracle.pyoracle import Oracle
    from oracle.urdf_model import load_urdf
    xml = data_urdf(urdf)
    m, om = Model.from_urdf(xml, armature=1e-2), load_urdf(xml, armature=1e-2)
    ev, orc = BatchEvaluator(m), Oracle(om)
    U, dt = 130, 0.05
    q, qd, tau, f, _ = random_inputs(om, U, seed=77)
    qd, tau = np.ascontiguousarray(8.0 * qd), np.ascontiguousarray(4.0 * tau)
    rq, rqd, rf, rj = orc.step_rk4_jvp(q, qd, tau, f, dt)
    d = [torch.from_numpy(a).cuda() for a in (q, qd, tau, f)]
    gq, gqd, gf, gj = ev.step_rk4_jvp(*d, dt)
    n = m.n
    assert rel_err_rows(gqd.cpu().numpy(), rqd) < TOL and rel_err_rows(gf.cpu().numpy(), rf) < TOL
    gj = gj.cpu().numpy()
    for r0, r1 in ((0, n), (n, 2 * n), (2 * n, 3 * n)):
        assert rel_err_rows(gj[r0:r1], rj[r0:r1]) < TOL, (r0, rel_err_rows(gj[r0:r1], rj[r0:r1]))
    Dq, Dv, M = ev.rnea_derivs(d[0], d[1])
    rDq, rDv, rM = orc.rnea_derivs(q, qd, None)
    for got, ref in ((Dq, rDq), (Dv, rDv), (M, rM)):
        assert rel_err_rows(got.cpu().numpy(), ref) < TOL


# ---------------------------------------------------------------------------------------------
# edge cases of the batch axis and of the workspace pipeline
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("U", [1, 31, 33, 96])
def test_ragged_batches_through_the_workspace_pipeline(U):
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import Model, data_urdf
    from oracle.pyoracle import Oracle
    from oracle.urdf_model import load_urdf
    xml = data_urdf("pilz6")
    m, om = Model.from_urdf(xml, armature=1e-2), load_urdf(xml, armature=1e-2)
    ev, orc = BatchEvaluator(m), Oracle(om)
    q, qd, tau, f, _ = random_inputs(om, U, seed=U)
    dtu = np.ascontiguousarray(np.random.default_rng(U).uniform(0.002, 0.03, U))
    rq, rqd, rf, rj = orc.step_rk4_jvp(q, qd, tau, f, 0.0, dt_u=dtu)
    d = [torch.from_numpy(a).cuda() for a in (q, qd, tau, f)]
    gq, gqd, gf, gj = ev.step_rk4_jvp(*d, torch.from_numpy(dtu).cuda())  # per-unit dt through K1/K3
    assert rel_err_rows(gqd.cpu().numpy(), rqd) < TOL and rel_err_rows(gf.cpu().numpy(), rf) < TOL
    assert rel_err_rows(gj.cpu().numpy(), rj) < TOL


def test_workspace_smaller_than_batch_is_chunked():
    """A caller-provided workspace for 64 units must serve U = 257 in five chunks with identical results."""
    import ctypes as C
    import torch
    from mpc_fatigue_b200 import _capi
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import Model, data_urdf
    from mpc_fatigue_b200.synth import synth_batch
    m = Model.from_urdf(data_urdf("pilz6"), armature=1e-2)
    ev = BatchEvaluator(m)
    lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
    U = 257
    q, qd, tau, f = synth_batch(lim, 0, U, 1, device="cuda")
    ref = ev.step_rk4_jvp(q, qd, tau, f, 0.02)
    need64 = int(_capi.lib.mpcf_step_rk4_jvp_workspace_bytes(m.handle, 64))
    assert need64 == 64 * 528 * 8 and int(_capi.lib.mpcf_step_rk4_jvp_workspace_bytes(m.handle, 1 << 24)) == (1 << 20) * 528 * 8
    ws = torch.empty(need64 // 8, dtype=torch.float64, device="cuda")
    out = [torch.empty_like(q) for _ in range(3)]
    jac = torch.empty((18, 25, U), dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    launches0 = _capi.lib.mpcf_launch_count()
    rc = _capi.lib.mpcf_step_rk4_jvp_ws_batch(m.handle, U, p(q), p(qd), p(tau), p(f), 0.02, None, p(out[0]), p(out[1]), p(out[2]),
                                              p(jac), p(ws), need64, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _capi.check(rc)
    assert _capi.lib.mpcf_launch_count() - launches0 == 3 * 5
    torch.cuda.synchronize()
    assert torch.equal(jac, ref[3]) and all(torch.equal(a, b) for a, b in zip(out, ref[:3]))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # a non-NULL workspace below one 32-unit tile, or a misaligned one, is an error (never a silent slow path)
    rc = _capi.lib.mpcf_step_rk4_jvp_ws_batch(m.handle, U, p(q), p(qd), p(tau), p(f), 0.02, None, p(out[0]), p(out[1]), p(out[2]),
                                              p(jac), p(ws), 1024, st)
    assert rc == _capi.EINVAL and "too small" in _capi.last_error()
    rc = _capi.lib.mpcf_step_rk4_jvp_ws_batch(m.handle, U, p(q), p(qd), p(tau), p(f), 0.02, None, p(out[0]), p(out[1]), p(out[2]),
                                              p(jac), C.c_void_p(ws.data_ptr() + 8), need64 - 8, st)
    assert rc == _capi.EINVAL and "aligned" in _capi.last_error()
    # the header-default entry (no workspace argument) runs the same pipeline on a pool-allocated workspace: bit-identical
    jac.zero_()
    launches0 = _capi.lib.mpcf_launch_count()
    _capi.check(_capi.lib.mpcf_step_rk4_jvp_batch(m.handle, U, p(q), p(qd), p(tau), p(f), 0.02, None, p(out[0]), p(out[1]), p(out[2]), p(jac), st))
    assert _capi.lib.mpcf_launch_count() - launches0 == 3
    torch.cuda.synchronize()
    assert torch.equal(jac, ref[3])
    # the dual-number sweeps agree to rounding
    _capi.check(_capi.lib.mpcf_step_rk4_jvp_dual_batch(m.handle, U, p(q), p(qd), p(tau), p(f), 0.02, None, p(out[0]), p(out[1]), p(out[2]), p(jac), st))
    assert float((jac - ref[3]).abs().max() / ref[3].abs().max()) < 1e-11


def test_max_dof_chain():
    """MPCF_MAX_DOF = 64 joints through the run-time-topology kernels (per-link state in local memory)."""
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import Model
    from oracle.pyoracle import Oracle
    m = Model.synthetic("chain", 64, seed=5, armature=1e-3)
    assert m.kernel_family == "generic64"
    om = oracle_model_from_export(m)
    ev, orc = BatchEvaluator(m), Oracle(om)
    U = 40
    q, qd, tau, f, qdd = random_inputs(om, U, seed=1)
    qd *= 0.1  # 64 stacked joints: keep the tip velocity moderate
    d = [torch.from_numpy(a).cuda() for a in (q, qd, tau, f, qdd)]
    assert rel_err_rows(ev.rnea(d[0], d[1], d[4]).cpu().numpy(), orc.rnea(q, qd, qdd)) < TOL
    assert rel_err_rows(ev.aba(d[0], d[1], d[2]).cpu().numpy(), orc.aba(q, qd, tau)) < 1e-8
    r = orc.step_rk4(q, qd, tau, f, 0.001)
    g = ev.step_rk4(d[0], d[1], d[2], d[3], 0.001)
    for a, b in zip(g, r):
        assert rel_err_rows(a.cpu().numpy(), b) < 1e-8


def test_zero_step_is_identity_with_k1_as_dt_column():
    """dt = 0: x+ = x, d x+/d x = I, d x+/d tau = 0, d x+/d dt = xdot(x) — exercises the dt column of the chain rule."""
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import Model, data_urdf
    from oracle.pyoracle import Oracle
    from oracle.urdf_model import load_urdf
    xml = data_urdf("pilz6")
    m, om = Model.from_urdf(xml, armature=1e-2), load_urdf(xml, armature=1e-2)
    ev, orc = BatchEvaluator(m), Oracle(om)
    U = 37
    q, qd, tau, f, _ = random_inputs(om, U, seed=8)
    d = [torch.from_numpy(a).cuda() for a in (q, qd, tau, f)]
    for direct in (False, True):
        gq, gqd, gf, gj = ev.step_rk4_jvp(*d, 0.0, direct=direct)
        assert torch.equal(gq, d[0]) and torch.equal(gqd, d[1]) and torch.equal(gf, d[3])
        J = gj.cpu().numpy()
        eye = np.zeros((18, 25))
        eye[:12, :12] = np.eye(12)
        eye[12:, 18:24] = np.eye(6)
        assert np.abs(J[:, :24] - eye[:, :24, None]).max() < 1e-15
        xdot = np.vstack([qd, orc.aba(q, qd, tau), orc.fatigue_rhs(f, tau, qd)])
        assert rel_err_rows(J[:, 24], xdot) < TOL
