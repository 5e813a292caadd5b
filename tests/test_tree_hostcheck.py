"""Host-side check of the run-time-tree Jacobian pipeline's arithmetic (mpc_fatigue_b200/csrc/tree_derivs.cuh and the column
recursion of kernels_tree.cu): tests/hostcheck/tree_host.cu compiles the same `__host__ __device__` source for the CPU, and the
results are compared with the oracle's complex-step derivatives — on the 37-joint branched tree of config C4 (prismatic +
revolute root chain, 6 limbs), on a chain and on the mixed-joint URDF.  Test infrastructure only: the product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, oracle_model_from_export, random_inputs
from mpc_fatigue_b200.model import Model
from oracle.pyoracle import Oracle

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(ROOT, "tests", "hostcheck", "tree_host.cu")
    out = os.path.join(ROOT, "tests", "hostcheck", "_build", "libtreehost.so")
    deps = [src] + [os.path.join(ROOT, "mpc_fatigue_b200", "csrc", f) for f in ("tree_derivs.cuh", "derivs.cuh", "dyn.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-x", "cu", "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-w",
                        "-I", os.path.join(ROOT, "mpc_fatigue_b200", "csrc"), src, "-o", out], check=True)
    return C.CDLL(out)


def _p(a, t=_DP):
    return a.ctypes.data_as(t)


def _model_args(m):
    a = {k: np.ascontiguousarray(m.export(k)) for k in ("parent", "jtype", "Rp", "pp", "mass", "mc", "Io", "arm", "grav", "fat")}
    return a, [m.n, _p(a["parent"], _IP), _p(a["jtype"], _IP)] + [_p(a[k]) for k in ("Rp", "pp", "mass", "mc", "Io", "arm", "grav")]


def _derivs(lib, m, margs, q, qd, qdd):
    n = m.n
    out = [np.zeros((n, n)) for _ in range(5)]
    rc = lib.hc_tree_derivs(*margs, _p(np.ascontiguousarray(q)), _p(np.ascontiguousarray(qd)), _p(np.ascontiguousarray(qdd)), *[_p(o) for o in out])
    assert rc == 0
    return out


MODELS = {
    "humanoid37": lambda: Model.synthetic("humanoid", 37, seed=7, armature=1e-2),
    "chain9": lambda: Model.synthetic("chain", 9, seed=4, armature=1e-2),
    "mixed": lambda: Model.from_urdf(open(os.path.join(ROOT, "tests", "golden", "mixed_joints.urdf")).read(), armature=1e-3),
}


@pytest.mark.parametrize("name", list(MODELS))
def test_tree_id_derivatives_and_factorisation(lib, name):
    m = MODELS[name]()
    om = oracle_model_from_export(m)
    orc = Oracle(om)
    n, U = m.n, 6
    q, qd, _, _, qdd = random_inputs(om, U, seed=31)
    rDq, rDv, rM = orc.rnea_derivs(q, qd, qdd)
    keep, margs = _model_args(m)
    par = keep["parent"]
    for u in range(U):
        Dq, Dv, M, Lf, Ci = _derivs(lib, m, margs, q[:, u], qd[:, u], qdd[:, u])
        assert np.abs(Ci @ M - np.eye(n)).max() < 1e-9  # M^-1 column by column from the packed factor
        for got, ref, nm in ((Dq, rDq, "dID/dq"), (Dv, rDv, "dID/dqd"), (M, rM, "M")):
            ref = ref[:, u].reshape(n, n)
            assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), (name, nm, u)
        # M = L^T D L with L unit lower triangular on the ancestor pattern
        L = np.tril(Lf, -1) + np.eye(n)
        D = np.diag(np.diag(Lf))
        assert np.abs(L.T @ D @ L - M).max() < 1e-11 * np.abs(M).max()
        # the in-place inverse of the unit factor (what the tensor-core chain kernel multiplies by): same pattern, L^-1 L = 1
        Li = np.zeros((n, n))
        assert lib.hc_tree_linv(*margs, _p(np.ascontiguousarray(q[:, u])), _p(np.ascontiguousarray(qd[:, u])), _p(np.ascontiguousarray(qdd[:, u])), _p(Li)) == 0
        assert np.array_equal(np.diag(Li), np.diag(Lf)) and np.array_equal(Li != 0, Lf != 0)
        Linv = np.tril(Li, -1) + np.eye(n)
        assert np.abs(Linv @ L - np.eye(n)).max() < 1e-12
        assert np.abs(Linv @ np.diag(1.0 / np.diag(Lf)) @ Linv.T @ M - np.eye(n)).max() < 1e-9
        for k in range(n):  # no fill-in outside the pattern
            anc, j = set(), par[k]
            while j >= 0:
                anc.add(int(j))
                j = par[j]
            assert all(Lf[k, j] == 0.0 for j in range(k) if j not in anc)


@pytest.mark.parametrize("name,dt", [("humanoid37", 0.0125), ("chain9", 0.02), ("mixed", 0.01), ("chain9", 0.0)])
def test_tree_column_recursion_gives_the_step_jacobian(lib, name, dt):
    m = MODELS[name]()
    om = oracle_model_from_export(m)
    orc = Oracle(om)
    n, U = m.n, 3
    q, qd, tau, f, _ = random_inputs(om, U, seed=32)
    _, _, _, rj = orc.step_rk4_jvp(q, qd, tau, f, dt)
    keep, margs = _model_args(m)
    fat = keep["fat"]
    # primal RK4 stages from the oracle's forward dynamics (the GPU takes them from k_tree_stages)
    cs = [0.5, 0.5, 1.0]
    xq, xv, xf = q.copy(), qd.copy(), f.copy()
    st = []
    for s in range(4):
        qdd = orc.aba(np.ascontiguousarray(xq), np.ascontiguousarray(xv), tau)
        fdot = orc.fatigue_rhs(np.ascontiguousarray(xf), tau, np.ascontiguousarray(xv))
        st.append((xq.copy(), xv.copy(), qdd, fdot))
        if s < 3:
            xq, xv, xf = q + cs[s] * dt * st[s][1], qd + cs[s] * dt * qdd, f + cs[s] * dt * fdot
    for u in range(U):
        mats = [_derivs(lib, m, margs, st[s][0][:, u], st[s][1][:, u], st[s][2][:, u]) for s in range(4)]
        Dq = np.ascontiguousarray(np.stack([mm[0] for mm in mats]))
        Dv = np.ascontiguousarray(np.stack([mm[1] for mm in mats]))
        Lf = np.ascontiguousarray(np.stack([mm[3] for mm in mats]))
        qds = np.ascontiguousarray(np.stack([st[s][1][:, u] for s in range(4)]))
        qdds = np.ascontiguousarray(np.stack([st[s][2][:, u] for s in range(4)]))
        fdots = np.ascontiguousarray(np.stack([st[s][3][:, u] for s in range(4)]))
        jac = np.zeros((3 * n, 4 * n + 1))
        lib.hc_tree_chain(n, _p(Dq), _p(Dv), _p(Lf), _p(qds), _p(qdds), _p(fdots), _p(np.ascontiguousarray(fat)), _p(np.ascontiguousarray(tau[:, u])),
                          C.c_double(dt), _p(jac))
        ref = rj[:, :, u]
        scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-6 * max(1.0, np.abs(ref).max()))
        assert (np.abs(jac - ref) / scale).max() < 1e-9, (name, u, float((np.abs(jac - ref) / scale).max()))
