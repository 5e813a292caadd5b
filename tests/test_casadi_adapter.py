"""mpc_fatigue_b200/casadi_adapter.py executed against tests/stubs/casadi (casadi is not installable in this image): the
callback protocol and the block-diagonal Jacobian assembly run for real; on the GPU box the assembled Jacobians are compared
with Function.jacobian() and with the oracle (the call it replaces: nlpsol(...) with CasADi-derived
derivatives, python/Pilz_6_DOF/force_optimization_pilz_6DOF.py:195-197)."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, random_inputs, rel_err_rows
from mpc_fatigue_b200.model import data_urdf


@pytest.fixture()
def adapter():
    stub_dir = os.path.join(ROOT, "tests", "stubs")
    sys.path.insert(0, stub_dir)
    sys.modules.pop("casadi", None)
    import mpc_fatigue_b200.casadi_adapter as ad
    ad = importlib.reload(ad)
    assert ad.HAVE_CASADI and ad.casadi.__version__.endswith("stub")
    yield ad
    sys.path.remove(stub_dir)
    sys.modules.pop("casadi", None)
    importlib.reload(ad)


def test_block_diagonal_assembly(adapter):
    N, n = 5, 3
    rng = np.random.default_rng(0)
    planes = rng.normal(size=(n * n, N))  # plane row*n + col, one value per node
    sp = adapter._block_diag_sparsity(N, n)
    assert (sp.size1(), sp.size2(), sp.nnz()) == (N * n, N * n, N * n * n)
    r, c = sp.get_triplet()
    assert all(ri // n == ci // n for ri, ci in zip(r, c))  # every non-zero sits in a diagonal block
    dm = adapter._block_diag_dm(planes, N, n)
    want = np.zeros((N * n, N * n))
    for k in range(N):
        want[k * n:(k + 1) * n, k * n:(k + 1) * n] = planes[:, k].reshape(n, n)
    assert np.array_equal(dm.full(), want)


def test_adapter_degrades_without_casadi():
    sys.modules.pop("casadi", None)
    import mpc_fatigue_b200.casadi_adapter as ad
    ad = importlib.reload(ad)
    if not ad.HAVE_CASADI:
        with pytest.raises(ImportError, match="casadi is not installed"):
            ad.make_dyn_fatigue_step_callback(data_urdf("pilz6"), 4)
    assert ad.IPOPT_OPTIONS["ipopt.hessian_approximation"] == "limited-memory"


@pytest.mark.gpu
def test_inverse_dynamics_callback_against_oracle(adapter):
    from oracle.pyoracle import Oracle
    from oracle.urdf_model import load_urdf
    xml, N = data_urdf("pilz6"), 7
    om = load_urdf(xml, armature=1e-2)
    orc = Oracle(om)
    q, qd, _, _, qdd = random_inputs(om, N, seed=3)
    cb = adapter.make_inverse_dynamics_callback(xml, N, armature=1e-2)
    assert cb.name_in() == ["q", "qdot", "qddot"] and cb.name_out() == ["tau"]
    vec = lambda a: a.T.reshape(-1)  # [n, N] -> vec([N, n])
    tau = cb(vec(q), vec(qd), vec(qdd))
    assert rel_err_rows(np.asarray(tau).reshape(N, 6).T, orc.rnea(q, qd, qdd)) < 1e-9
    jac = cb.jacobian()
    assert jac.name_out() == ["jac_tau_q", "jac_tau_qdot", "jac_tau_qddot"]
    Jq, Jv, M = jac(vec(q), vec(qd), vec(qdd), np.zeros(6 * N))
    # the oracle's complex-step derivatives of RNEA, node by node
    Dq, Dv, Mo = orc.rnea_derivs(q, qd, qdd)
    for name, got, ref in (("q", Jq, Dq), ("qd", Jv, Dv), ("qdd", M, Mo)):
        full = got.full()
        for k in range(N):
            blk = full[6 * k:6 * k + 6, 6 * k:6 * k + 6]
            assert np.abs(blk - ref[:, k].reshape(6, 6)).max() < 1e-9 * max(1.0, np.abs(ref[:, k]).max()), (name, k)
            full[6 * k:6 * k + 6, 6 * k:6 * k + 6] = 0.0
        assert not full.any()  # nothing outside the diagonal blocks


@pytest.mark.gpu
def test_dyn_fatigue_step_callback_against_function_and_oracle(adapter):
    from mpc_fatigue_b200.pynocchio_casadi import Function, generate_fwd_dyn_fatigue_step
    from oracle.pyoracle import Oracle
    from oracle.urdf_model import load_urdf
    xml, N, dt, n = data_urdf("pilz6"), 9, 0.02, 6
    om = load_urdf(xml, armature=1e-2)
    orc = Oracle(om)
    q, qd, tau, f, _ = random_inputs(om, N, seed=4)
    cb = adapter.make_dyn_fatigue_step_callback(xml, N, armature=1e-2)
    assert cb.name_in() == ["q", "qd", "tau", "f", "dt"] and cb.name_out() == ["q_next", "qd_next", "f_next"]
    vec = lambda a: a.T.reshape(-1)
    outs = cb(vec(q), vec(qd), vec(tau), vec(f), dt)
    rq, rqd, rf, rj = orc.step_rk4_jvp(q, qd, tau, f, dt)
    for got, ref in zip(outs, (rq, rqd, rf)):
        assert rel_err_rows(np.asarray(got).reshape(N, n).T, ref) < 1e-9
    jac = cb.jacobian()
    assert jac.n_in() == 8 and jac.n_out() == 15 and jac.name_out(0) == "jac_q_next_q" and jac.name_out(14) == "jac_f_next_dt"
    blocks = jac(vec(q), vec(qd), vec(tau), vec(f), dt, *[np.zeros(n * N)] * 3)
    # reference 1: the Function look-alike's dense per-node Jacobian; reference 2: the oracle (complex step)
    step = Function.deserialize(generate_fwd_dyn_fatigue_step(xml, {"armature": 1e-2}))
    Jfn = step.jacobian()(q.T, qd.T, tau.T, f.T, dt)  # [N, 3n, 4n+1]
    for o in range(3):
        for i in range(5):
            B = blocks[o * 5 + i].full()
            for k in range(N):
                if i < 4:
                    blk = B[n * k:n * k + n, n * k:n * k + n].copy()
                    B[n * k:n * k + n, n * k:n * k + n] = 0.0
                    cols = slice(i * n, (i + 1) * n)
                else:
                    blk = B[n * k:n * k + n, 0:1]
                    cols = slice(4 * n, 4 * n + 1)
                ref_fn = Jfn[k, o * n:(o + 1) * n, cols]
                ref_or = rj[o * n:(o + 1) * n, cols, k]
                assert np.array_equal(blk, ref_fn), (o, i, k)
                assert np.abs(blk - ref_or).max() < 1e-9 * max(1.0, np.abs(ref_or).max()), (o, i, k)
            if i < 4:
                assert not B.any()
