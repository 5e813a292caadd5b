"""oracle/pyoracle.py — TEST INFRASTRUCTURE ONLY.

ctypes front-end of the C oracle (oracle/mpcf_oracle.c).  Arrays are numpy float64 in the SoA
layout of the C-ABI: `[component, U]`, C-contiguous.  Importers: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / `--impl reference` legs.  Nothing under mpc_fatigue_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .urdf_model import Model

_HERE = os.path.dirname(os.path.abspath(__file__))
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


class _CModel(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("nframes", C.c_int), ("parent", _IP), ("jtype", _IP), ("Rp", _DP), ("pp", _DP),
        ("mass", _DP), ("mc", _DP), ("Io", _DP), ("arm", _DP), ("fat", _DP), ("grav", C.c_double * 3),
        ("fparent", _IP), ("fR", _DP), ("fp", _DP),
    ]


def build(force: bool = False) -> None:
    """Compile the oracle libraries (checker + timing build) with the committed Makefile."""
    args = ["make", "-C", _HERE, "-s"] + (["-B"] if force else [])
    subprocess.run(args, check=True)


def _load(fast: bool):
    name = "libmpcf_oracle_fast.so" if fast else "libmpcf_oracle.so"
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.mpcfo_set_threads.restype = C.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_DP)


def _chk(a, rows, U):
    assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (rows, U), (a.shape, (rows, U))
    return a


class Oracle:
    """CPU oracle bound to one model.  `fast=True` selects the -O3 -march=native timing build."""

    def __init__(self, model: Model, fast: bool = False, threads: int = 0):
        self.lib = _load(fast)
        self.model = model
        self.n = model.n
        self._a = a = model.arrays()
        cm = _CModel()
        cm.n, cm.nframes = model.n, len(model.fparent)
        for k in ("parent", "jtype", "fparent"):
            setattr(cm, k, a[k].ctypes.data_as(_IP))
        for k in ("Rp", "pp", "mass", "mc", "Io", "arm", "fat", "fR", "fp"):
            setattr(cm, k, a[k].ctypes.data_as(_DP))
        cm.grav = (C.c_double * 3)(*model.grav)
        self._cm = cm
        self.threads = self.lib.mpcfo_set_threads(int(threads))

    def _ref(self):
        return C.byref(self._cm)

    def rnea(self, q, qd, qdd=None):
        n, U = q.shape
        tau = np.empty((n, U))
        rc = self.lib.mpcfo_rnea_batch(self._ref(), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)),
                                       _p(qdd if qdd is None else _chk(qdd, n, U)), _p(tau))
        assert rc == 0, rc
        return tau

    def fk(self, frame: int, q):
        n, U = q.shape
        pos, rot = np.empty((3, U)), np.empty((9, U))
        rc = self.lib.mpcfo_fk_batch(self._ref(), frame, C.c_long(U), _p(_chk(q, n, U)), _p(pos), _p(rot))
        assert rc == 0, rc
        return pos, rot

    def jacobian(self, frame: int, q):
        n, U = q.shape
        J = np.empty((6 * n, U))
        rc = self.lib.mpcfo_jac_batch(self._ref(), frame, C.c_long(U), _p(_chk(q, n, U)), _p(J))
        assert rc == 0, rc
        return J

    def crba(self, q1):
        q1 = np.ascontiguousarray(q1, dtype=np.float64)
        M = np.empty((self.n, self.n))
        rc = self.lib.mpcfo_crba(self._ref(), _p(q1), _p(M))
        assert rc == 0, rc
        return M

    def aba(self, q, qd, tau):
        n, U = q.shape
        qdd = np.empty((n, U))
        rc = self.lib.mpcfo_aba_batch(self._ref(), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)),
                                      _p(_chk(tau, n, U)), _p(qdd))
        if rc != 0:
            raise ZeroDivisionError("ABA singular (zero joint-space inertia without armature)")
        return qdd

    def node_eval_ref(self, ee_frames, wsign, q, qd, W, T, h, qdd=None):
        n, U = q.shape
        nee = len(ee_frames)
        fr = (C.c_int * max(nee, 1))(*ee_frames)
        tau, qn, Tn = np.empty((n, U)), np.empty((n, U)), np.empty((n, U))
        if W is None:
            W = np.zeros((6 * max(nee, 1), U))
        rc = self.lib.mpcfo_node_eval_ref_batch(
            self._ref(), nee, fr, C.c_double(wsign), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)),
            _p(qdd), _p(W), _p(T), C.c_double(h), _p(tau), _p(qn), _p(Tn))
        assert rc == 0, rc
        return tau, qn, Tn

    def step_rk4(self, q, qd, tau, f, dt, dt_u=None):
        n, U = q.shape
        qn, qdn, fn = np.empty((n, U)), np.empty((n, U)), np.empty((n, U))
        rc = self.lib.mpcfo_step_rk4_batch(self._ref(), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)),
                                           _p(_chk(tau, n, U)), _p(_chk(f, n, U)), C.c_double(dt), _p(dt_u),
                                           _p(qn), _p(qdn), _p(fn))
        if rc != 0:
            raise ZeroDivisionError("ABA singular")
        return qn, qdn, fn

    def step_rk4_jvp(self, q, qd, tau, f, dt, dt_u=None):
        n, U = q.shape
        qn, qdn, fn = np.empty((n, U)), np.empty((n, U)), np.empty((n, U))
        jac = np.empty((3 * n, 4 * n + 1, U))
        rc = self.lib.mpcfo_step_rk4_jvp_batch(self._ref(), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)),
                                               _p(_chk(tau, n, U)), _p(_chk(f, n, U)), C.c_double(dt), _p(dt_u),
                                               _p(qn), _p(qdn), _p(fn), _p(jac))
        if rc != 0:
            raise ZeroDivisionError("ABA singular")
        return qn, qdn, fn, jac

    def _coupling(self, ee_frames, weight):
        n = self.n
        par = self._a["parent"]
        root = list(range(n))
        for i in range(n):
            root[i] = i if par[i] < 0 else root[par[i]]
        roots = sorted(set(root))
        assert len(roots) == 2, "coupled fatigue needs a two-arm model"
        chain_of = np.ascontiguousarray(np.array([roots.index(r) for r in root], dtype=np.int32))

        class _Cpl(C.Structure):
            _fields_ = [("chain_of", _IP), ("ee_frame", C.c_int * 2), ("weight", C.c_double)]
        cp = _Cpl()
        cp.chain_of = chain_of.ctypes.data_as(_IP)
        cp.ee_frame = (C.c_int * 2)(*[int(x) for x in ee_frames])
        cp.weight = float(weight)
        cp._keep = chain_of
        return cp

    def step_rk4_coupled(self, ee_frames, weight, q, qd, tau, f, dt, dt_u=None, jac=False):
        """Coupled-fatigue RK4 step of a two-arm model (core.inc.h: step_rk4_coupled); jac=True adds the complex-step Jacobian."""
        n, U = q.shape
        cp = self._coupling(ee_frames, weight)
        qn, qdn, fn = np.empty((n, U)), np.empty((n, U)), np.empty((n, U))
        args = [self._ref(), C.byref(cp), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)), _p(_chk(tau, n, U)), _p(_chk(f, n, U)),
                C.c_double(dt), _p(dt_u), _p(qn), _p(qdn), _p(fn)]
        if jac:
            J = np.empty((3 * n, 4 * n + 1, U))
            rc = self.lib.mpcfo_step_rk4_coupled_jvp_batch(*args, _p(J))
        else:
            rc = self.lib.mpcfo_step_rk4_coupled_batch(*args)
        if rc != 0:
            raise ZeroDivisionError("ABA singular")
        return (qn, qdn, fn, J) if jac else (qn, qdn, fn)

    def ocp_rows(self, opts: dict, B, N, q, qd, F, T=None, q_last=None, T_last=None, rel_pos0=None, rel_ori0=None, kin_jac=False):
        """Reference-mode OCP node rows (core.inc.h: ocp_rows).  opts: narm, ee_frame (list), wsign, fdes, dist2_ref, mu, p_ref,
        w_box, w_qd, w_F, h.  Returns rows [nrows, U], cost [U] (and kin_jac [kin, n + 3 narm, U] by complex step)."""
        class _RO(C.Structure):
            _fields_ = [("narm", C.c_int), ("ee_frame", C.c_int * 2), ("wsign", C.c_double), ("fdes", C.c_double * 3), ("dist2_ref", C.c_double),
                        ("mu", C.c_double), ("p_ref", C.c_double * 3), ("w_box", C.c_double), ("w_qd", C.c_double), ("w_F", C.c_double), ("h", C.c_double)]
        narm = int(opts["narm"])
        ee = list(opts["ee_frame"]) + [0] * (2 - len(opts["ee_frame"]))
        o = _RO(narm, (C.c_int * 2)(*ee), opts.get("wsign", -1.0), (C.c_double * 3)(*opts.get("fdes", (0, 0, 0))), opts.get("dist2_ref", 0.0),
                opts.get("mu", 0.0), (C.c_double * 3)(*opts.get("p_ref", (0, 0, 0))), opts.get("w_box", 0.0), opts.get("w_qd", 0.0), opts.get("w_F", 0.0),
                opts.get("h", 0.0))
        n, U = q.shape
        assert U == B * N
        kin = 26 if narm == 2 else 3
        rows, cost = np.empty((kin + 3 * n, U)), np.empty(U)
        J = np.empty((kin, n + 3 * narm, U)) if kin_jac else None
        rc = self.lib.mpcfo_ocp_rows_batch(self._ref(), C.byref(o), C.c_long(B), C.c_int(N), _p(_chk(q, n, U)), _p(_chk(qd, n, U)), _p(_chk(F, 3 * narm, U)),
                                           _p(T), _p(q_last), _p(T_last), _p(rel_pos0), _p(rel_ori0), _p(rows), _p(cost), _p(J))
        assert rc == 0, rc
        return (rows, cost, J) if kin_jac else (rows, cost)

    def step_rk4_jvp_forward(self, q, qd, tau, f, dt, dt_u=None):
        """Same result as step_rk4_jvp by ONE forward-mode sweep per unit with all 3n + 1 directions in SIMD lanes
        (oracle/forward_mode.cpp, -O3 -march=native, OpenMP): the CPU baseline of bench.py.  n <= 7."""
        lib = getattr(self, "_fwd", None)
        if lib is None:
            path = os.path.join(_HERE, "_build", "libmpcf_oracle_fwd.so")
            if not os.path.exists(path):
                build()
            lib = self._fwd = C.CDLL(path)
            lib.mpcfo_fwd_set_threads(int(self.threads))
        n, U = q.shape
        P = 4 * n + 1
        qn, qdn, fn = np.empty((n, U)), np.empty((n, U)), np.empty((n, U))
        jac = np.empty((3 * n, P, U))
        rc = lib.mpcfo_step_rk4_jvp_forward_batch(self._ref(), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)),
                                                  _p(_chk(tau, n, U)), _p(_chk(f, n, U)), C.c_double(dt), _p(dt_u),
                                                  _p(qn), _p(qdn), _p(fn), _p(jac))
        if rc == -1:
            raise ValueError("forward-mode baseline supports n <= 7")
        if rc != 0:
            raise ZeroDivisionError("ABA singular")
        return qn, qdn, fn, jac

    def fatigue_zoh(self, T, tau, qd, h):
        n, U = T.shape
        out = np.empty((n, U))
        rc = self.lib.mpcfo_fatigue_zoh_batch(self._ref(), C.c_long(U), _p(_chk(T, n, U)), _p(_chk(tau, n, U)),
                                              _p(_chk(qd, n, U)), C.c_double(h), _p(out))
        assert rc == 0, rc
        return out

    def fatigue_rhs(self, f, tau, qd):
        n, U = f.shape
        out = np.empty((n, U))
        rc = self.lib.mpcfo_fatigue_rhs_batch(self._ref(), C.c_long(U), _p(_chk(f, n, U)), _p(_chk(tau, n, U)),
                                              _p(_chk(qd, n, U)), _p(out))
        assert rc == 0, rc
        return out

    def fd_derivs(self, q, qd, tau):
        """A = d qdd/d q, B = d qdd/d qd, C = M^-1 as [n*n, U] planes (row*n + col)."""
        n, U = q.shape
        A, B, Cm = np.empty((n * n, U)), np.empty((n * n, U)), np.empty((n * n, U))
        rc = self.lib.mpcfo_fd_derivs_batch(self._ref(), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)),
                                            _p(_chk(tau, n, U)), _p(A), _p(B), _p(Cm))
        if rc != 0:
            raise ZeroDivisionError("ABA singular")
        return A, B, Cm

    def rnea_derivs(self, q, qd, qdd=None):
        """d tau/d q, d tau/d qd, M = d tau/d qdd as [n*n, U] planes (row*n + col)."""
        n, U = q.shape
        Dq, Dv, M = np.empty((n * n, U)), np.empty((n * n, U)), np.empty((n * n, U))
        rc = self.lib.mpcfo_rnea_derivs_batch(self._ref(), C.c_long(U), _p(_chk(q, n, U)), _p(_chk(qd, n, U)),
                                              _p(qdd if qdd is None else _chk(qdd, n, U)), _p(Dq), _p(Dv), _p(M))
        assert rc == 0, rc
        return Dq, Dv, M

    def node_eval_ref_jvp(self, ee_frames, wsign, q, qd, W, qdd=None):
        """d tau/d q, d tau/d qd of tau = RNEA + wsign * sum J^T W as [n*n, U] planes."""
        n, U = q.shape
        nee = len(ee_frames)
        fr = (C.c_int * max(nee, 1))(*ee_frames)
        Dq, Dv = np.empty((n * n, U)), np.empty((n * n, U))
        if W is None:
            W = np.zeros((6 * max(nee, 1), U))
        rc = self.lib.mpcfo_node_eval_ref_jvp_batch(self._ref(), nee, fr, C.c_double(wsign), C.c_long(U), _p(_chk(q, n, U)),
                                                    _p(_chk(qd, n, U)), _p(qdd), _p(W), _p(Dq), _p(Dv))
        assert rc == 0, rc
        return Dq, Dv
