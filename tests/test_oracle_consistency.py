"""Self-consistency of the oracle for everything the reference never computes (ABA, RK4, Jacobians) —
"parity unpinned" items are pinned here by independent identities (SURVEY.md §4 "what nothing pins")."""
import numpy as np
import pytest

from conftest import oracle_model_from_export, random_inputs, rel_err_rows
from mpc_fatigue_b200.model import Model, data_urdf
from oracle.pyoracle import Oracle
from oracle.urdf_model import load_urdf

MODELS = ["pilz6", "pilz3", "pilz6x2", "humanoid37"]


@pytest.fixture(scope="module", params=MODELS)
def case(request):
    if request.param == "humanoid37":
        om = oracle_model_from_export(Model.synthetic("humanoid", 37, seed=7, armature=1e-2))
    else:
        om = load_urdf(data_urdf(request.param), armature=1e-2)
    U = 24 if om.n <= 12 else 6
    return om, Oracle(om), random_inputs(om, U, seed=3), U


def test_rnea_of_aba_is_identity(case):
    om, orc, (q, qd, tau, f, qdd), U = case
    a = orc.aba(q, qd, tau)
    assert np.abs(orc.rnea(q, qd, a) - tau).max() < 1e-9 * max(1.0, np.abs(tau).max())


def test_crba_matches_rnea_columns_and_aba(case):
    om, orc, (q, qd, tau, f, qdd), U = case
    n = om.n
    for u in range(min(U, 4)):
        qu, qdu, tu = (np.ascontiguousarray(a[:, u:u + 1]) for a in (q, qd, tau))
        M = orc.crba(qu[:, 0].copy())
        z = np.zeros((n, 1))
        h0 = orc.rnea(qu, z, z)
        cols = np.hstack([orc.rnea(qu, z, np.eye(n)[:, j:j + 1].copy()) - h0 for j in range(n)])
        assert np.abs(M - cols).max() < 1e-12 * max(1.0, np.abs(M).max())
        assert np.abs(M - M.T).max() == 0.0
        assert np.linalg.eigvalsh(M).min() > 0
        h = orc.rnea(qu, qdu, z)
        ref = np.linalg.solve(M, tu - h)
        assert np.abs(orc.aba(qu, qdu, tu) - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_aba_singular_without_armature():
    orc = Oracle(load_urdf(data_urdf("pilz6"), armature=0.0))  # zero-inertia flange: M[5,5] = 0 (SURVEY.md §0 trap)
    z = np.zeros((6, 1))
    with pytest.raises(ZeroDivisionError):
        orc.aba(z, z, z)


def test_jvp_vs_central_differences(case):
    om, orc, (q, qd, tau, f, qdd), U = case
    n, dt = om.n, 0.02
    U = min(U, 4)
    q, qd, tau, f = (np.ascontiguousarray(a[:, :U]) for a in (q, qd, tau, f))
    qn, qdn, fn, jac = orc.step_rk4_jvp(q, qd, tau, f, dt)
    p = orc.step_rk4(q, qd, tau, f, dt)
    assert max(np.abs(a - b).max() for a, b in zip((qn, qdn, fn), p)) < 1e-13 * max(1.0, np.abs(qdn).max())
    X = np.vstack([q, qd, tau, f])
    eps = 1e-6
    dirs = range(4 * n + 1) if n <= 12 else list(range(0, 4 * n, 9)) + [4 * n]
    for d in dirs:
        def run(sg):
            Xp, dtp = X.copy(), dt
            if d < 4 * n:
                Xp[d] += sg * eps
            else:
                dtp += sg * eps
            a, b, c = orc.step_rk4(*(np.ascontiguousarray(Xp[k * n:(k + 1) * n]) for k in range(4)), dtp)
            return np.vstack([a, b, c])
        fd = (run(1) - run(-1)) / (2 * eps)
        scale = max(1.0, np.abs(jac[:, d]).max())
        assert np.abs(fd - jac[:, d]).max() < 2e-6 * scale, d


def test_jacobian_structure(case):
    om, orc, (q, qd, tau, f, qdd), U = case
    n = om.n
    U = min(U, 3)
    _, _, _, jac = orc.step_rk4_jvp(*(np.ascontiguousarray(a[:, :U]) for a in (q, qd, tau, f)), 0.02)
    Jf = jac[:, 3 * n:4 * n]
    assert np.all(Jf[:2 * n] == 0.0)  # (q+, qd+) do not depend on f
    assert np.all(Jf[2 * n:] * (1 - np.eye(n))[:, :, None] == 0.0)  # df+/df diagonal
    lam, z = np.array(om.fat)[:, 0], None
    z = lam * 0.02
    g = 1 - z + z ** 2 / 2 - z ** 3 / 6 + z ** 4 / 24
    assert np.abs(np.einsum("iiu->iu", Jf[2 * n:]) - g[:, None]).max() < 1e-14


def test_rk4_order_and_thermal_zoh():
    om = load_urdf(data_urdf("pilz6"), armature=1e-2)
    orc = Oracle(om)
    q, qd, tau, f, _ = random_inputs(om, 4, seed=9)
    qd *= 0.2
    T = 0.02

    def integrate(nsteps):
        x = (q, qd, f)
        for _ in range(nsteps):
            x = orc.step_rk4(x[0], x[1], tau, x[2], T / nsteps)
        return np.vstack(x)
    ref = integrate(32)
    e1, e2 = np.abs(integrate(1) - ref).max(), np.abs(integrate(2) - ref).max()
    assert 8.0 < e1 / e2 < 40.0  # 4th order: error ratio ~ 16
    # fatigue alone, constant (tau, qd): RK4 differs from the exact ZOH map by ~ (h/tau_theta)^5 / 120
    om2 = load_urdf(data_urdf("pilz3"))
    o2 = Oracle(om2)
    Tw, tt, z = np.full((3, 1), 40.0), np.full((3, 1), 30.0), np.zeros((3, 1))
    h = 0.5
    exact = o2.fatigue_zoh(Tw, tt, z, h)
    lam, kap, ct, cv = om2.fat[0]
    u = kap * ct * 30.0 ** 2

    def rhs(x):
        return -lam * x + u
    k1 = rhs(40.0); k2 = rhs(40.0 + h / 2 * k1); k3 = rhs(40.0 + h / 2 * k2); k4 = rhs(40.0 + h * k3)
    rk = 40.0 + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    assert abs(rk - exact[0, 0]) < 1e-12
    assert np.abs(o2.fatigue_rhs(Tw, tt, z) - rhs(40.0)).max() < 1e-15


def test_gravity_only_static_torque_is_configuration_gradient_of_potential():
    """tau(q, 0, 0) = dV/dq with V = -sum m_i g . c_i(q): checks frames, COM handling and gravity sign by finite differences."""
    om = load_urdf(data_urdf("pilz6"))
    orc = Oracle(om)
    a = om.arrays()
    rng = np.random.default_rng(4)
    q = rng.uniform(-2, 2, (6, 1))

    def potential(qv):
        # world COM of each body through the oracle's own FK of the joint frames
        V = 0.0
        for i in range(6):
            fr = om.frame_id(om.joint_names[i])
            p, R = orc.fk(fr, np.ascontiguousarray(qv))
            if a["mass"][i] > 0:
                c = a["mc"][i] / a["mass"][i]
                V += a["mass"][i] * 9.81 * (p[2, 0] + R[6:9, 0] @ c)
        return V
    tau = orc.rnea(q, np.zeros((6, 1)), np.zeros((6, 1)))[:, 0]
    g = np.zeros(6)
    for j in range(6):
        e = np.zeros((6, 1)); e[j] = 1e-6
        g[j] = (potential(q + e) - potential(q - e)) / 2e-6
    assert np.abs(tau - g).max() < 1e-6


def test_forward_mode_baseline_matches_the_complex_step_checker():
    """oracle/forward_mode.cpp (bench.py's CPU baseline: one forward-mode sweep, directions in SIMD lanes, closed-form fatigue
    columns) against the complex-step instantiation, on the three models it supports (n <= 7)."""
    from mpc_fatigue_b200.model import data_urdf
    from oracle.urdf_model import load_urdf
    for name, U in (("pilz6", 257), ("pilz3", 64)):
        om = load_urdf(data_urdf(name), armature=1e-2)
        orc = Oracle(om)
        q, qd, tau, f, _ = random_inputs(om, U, seed=17)
        dt_u = np.ascontiguousarray(np.random.default_rng(3).uniform(0.005, 0.03, U))
        for kw in (dict(dt_u=None), dict(dt_u=dt_u)):
            ref = orc.step_rk4_jvp(q, qd, tau, f, 0.02, **kw)
            got = orc.step_rk4_jvp_forward(q, qd, tau, f, 0.02, **kw)
            for a, b in zip(got[:3], ref[:3]):
                assert rel_err_rows(a, b) < 1e-13
            n = om.n
            assert rel_err_rows(got[3].reshape(3 * n * (4 * n + 1), U), ref[3].reshape(3 * n * (4 * n + 1), U)) < 1e-10
