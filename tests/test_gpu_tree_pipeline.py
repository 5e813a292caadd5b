"""GPU: the run-time-tree Jacobian pipeline (csrc/kernels_tree.cu: k_tree_stages -> k_tree_derivs -> k_tree_chain_tc) against the
oracle's complex-step Jacobian, per Jacobian plane, on the 37-joint branched tree of config C4 (prismatic + revolute root
chain, six limbs), a short chain and the mixed-joint URDF; ragged batches, per-unit dt, dt = 0, workspace chunking, and
agreement with the dual-number sweeps it replaces."""
import os

import numpy as np
import pytest

from conftest import ROOT, oracle_model_from_export, random_inputs, rel_err_rows
from mpc_fatigue_b200.model import Model
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
TOL = 1e-9

MODELS = {
    "humanoid37": lambda: Model.synthetic("humanoid", 37, seed=7, armature=1e-2),
    "chain9": lambda: Model.synthetic("chain", 9, seed=4, armature=1e-2),
    "chain40": lambda: Model.synthetic("chain", 40, seed=6, armature=1e-2),      # two full 64-column slabs (121 columns)
    "humanoid21": lambda: Model.synthetic("humanoid", 21, seed=3, armature=1e-2),  # one slab of the 40-row kernel (64 columns)
    "humanoid25": lambda: Model.synthetic("humanoid", 25, seed=5, armature=1e-2),  # 76 columns: a second, mostly empty slab
    "two_roots10": lambda: Model.synthetic("dual_arm", 10, seed=3, armature=1e-2),  # two trees in one model (16-row kernel)
    "two_roots18": lambda: Model.synthetic("dual_arm", 18, seed=3, armature=1e-2),  # the same through the 40-row kernel
    "mixed": lambda: Model.from_urdf(open(os.path.join(ROOT, "tests", "golden", "mixed_joints.urdf")).read(), armature=1e-3),
}


def _check(m, U, dt, seed, dt_u=None):
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    om = oracle_model_from_export(m)
    orc, ev = Oracle(om, fast=True), BatchEvaluator(m)
    n = m.n
    q, qd, tau, f, _ = random_inputs(om, U, seed=seed)
    d = [torch.from_numpy(a).cuda() for a in (q, qd, tau, f)]
    dtg = dt if dt_u is None else torch.from_numpy(dt_u).cuda()
    rq, rqd, rf, rj = orc.step_rk4_jvp(q, qd, tau, f, dt, dt_u=dt_u)
    gq, gqd, gf, gj = [t.cpu().numpy() for t in ev.step_rk4_jvp(*d, dtg)]
    assert rel_err_rows(gq, rq) < TOL and rel_err_rows(gqd, rqd) < TOL and rel_err_rows(gf, rf) < TOL
    P = 4 * n + 1
    err = rel_err_rows(gj.reshape(3 * n * P, U), rj.reshape(3 * n * P, U))
    assert err < TOL, err
    return ev, d, gj


@pytest.mark.parametrize("name", list(MODELS))
def test_tree_pipeline_matches_the_oracle(name):
    m = MODELS[name]()
    assert m.kernel_family.startswith("generic")
    _check(m, 97 if name != "chain40" else 33, 0.0125, seed=41)


@pytest.mark.parametrize("U", [1, 31, 33, 130])
def test_tree_pipeline_ragged_batches(U):
    _check(MODELS["humanoid37"](), U, 0.0125, seed=42)


def test_tree_pipeline_per_unit_dt_and_zero_dt():
    m = MODELS["humanoid37"]()
    dt_u = np.ascontiguousarray(np.random.default_rng(1).uniform(0.002, 0.02, 65))
    _check(m, 65, 0.0, seed=43, dt_u=dt_u)
    _check(m, 40, 0.0, seed=44)  # dt = 0: identity / k1 columns, no division by h anywhere


def test_tree_pipeline_agrees_with_the_dual_number_sweeps_and_chunks():
    import ctypes as C
    import torch
    from mpc_fatigue_b200 import _capi
    m = MODELS["humanoid37"]()
    ev, d, gj = _check(m, 200, 0.0125, seed=45)
    dual = ev.step_rk4_jvp(*d, 0.0125, direct=True)[3].cpu().numpy()
    assert float(np.abs(dual - gj).max() / np.abs(dual).max()) < 1e-10
    # a caller workspace that holds 64 units serves U = 200 in four chunks, bit-identical
    need = int(_capi.lib.mpcf_step_rk4_jvp_workspace_bytes(m.handle, 64))
    full = int(_capi.lib.mpcf_step_rk4_jvp_workspace_bytes(m.handle, 200))
    assert need < full
    ws = torch.empty(need // 8, dtype=torch.float64, device="cuda")
    out = [torch.empty_like(d[0]) for _ in range(3)]
    jac = torch.empty((111, 149, 200), dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    _capi.check(_capi.lib.mpcf_step_rk4_jvp_ws_batch(m.handle, 200, p(d[0]), p(d[1]), p(d[2]), p(d[3]), 0.0125, None, p(out[0]), p(out[1]), p(out[2]),
                                                     p(jac), p(ws), need, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert np.array_equal(jac.cpu().numpy(), gj)


def test_c4_sample_of_the_full_size_batch():
    """4,096 units drawn from the 32,768 x 40 batch of config C4 (same generator, same seed as bench.py), oracle-checked."""
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.synth import synth_batch
    m = MODELS["humanoid37"]()
    ev = BatchEvaluator(m)
    om = oracle_model_from_export(m)
    orc = Oracle(om, fast=True)
    B, N, n = 32768, 40, 37
    lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
    rng = np.random.default_rng(7)
    scen = np.sort(rng.choice(B, 128, replace=False))
    # scenario i's inputs are a pure function of (seed, i): generate the 128 sampled scenarios x 40 nodes = 5,120 units
    parts = [synth_batch(lim, int(b), 1, N, seed=1234, device="cuda") for b in scen]
    q, qd, tau, f = (torch.cat([p[k] for p in parts], dim=1).contiguous() for k in range(4))
    U = q.shape[1]
    assert U == 128 * N
    gq, gqd, gf, gj = [t.cpu().numpy() for t in ev.step_rk4_jvp(q, qd, tau, f, 0.5 / 40)]
    rq, rqd, rf, rj = orc.step_rk4_jvp(*[t.cpu().numpy() for t in (q, qd, tau, f)], 0.5 / 40)
    assert rel_err_rows(gq, rq) < TOL and rel_err_rows(gqd, rqd) < TOL and rel_err_rows(gf, rf) < TOL
    assert rel_err_rows(gj.reshape(-1, U), rj.reshape(-1, U)) < TOL


@pytest.mark.parametrize("name", ["humanoid37", "chain9", "chain40", "mixed"])
def test_scalar_chain_kernel_matches_the_oracle(monkeypatch, name):
    """MPCF_TREE_CHAIN=scalar: the DFMA chain kernel (triangular solves with L instead of the tensor-core products with L^-1) is
    kept as an independent cross-check of the default path; both must agree with the oracle."""
    monkeypatch.setenv("MPCF_TREE_CHAIN", "scalar")
    _check(MODELS[name](), 33 if name == "chain40" else 130, 0.0125, seed=46)
    if name == "humanoid37":
        dt_u = np.ascontiguousarray(np.random.default_rng(2).uniform(0.002, 0.02, 33))
        _check(MODELS[name](), 33, 0.0, seed=47, dt_u=dt_u)
