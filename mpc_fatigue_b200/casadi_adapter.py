"""casadi.Callback adapters: put the GPU evaluator inside a CasADi MX graph / an IPOPT solve.

The reference builds its NLPs from SX-traced Functions and lets CasADi differentiate the graph
(`nlpsol('solver', 'ipopt', nlp)`, python/Pilz_6_DOF/force_optimization_pilz_6DOF.py:195-197).  A numeric GPU
evaluator cannot be traced by SX; it enters an MX graph as a `casadi.Callback` that supplies its own Jacobian, and
IPOPT runs with `hessian_approximation = limited-memory`.

casadi is NOT installed in the build image, so this module is import-guarded.  Its logic (callback protocol, block-diagonal
Jacobian assembly, input/output stacking) is executed in the tests against a minimal stand-in of the casadi classes it
touches (tests/stubs/casadi: Sparsity.dense/triplet, DM, Callback with the CasADi >= 3.6 `get_jacobian` convention: inputs =
inputs + nominal outputs, outputs = one block `jac_<out>_<in>` per pair), and the assembled blocks are checked against
`Function.jacobian()` and the oracle.  Layouts follow CasADi: every input and output is a dense column vector; batched
quantities are stacked node-major, i.e. `q = vec([N, n])`.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - casadi is absent in the build image
    import casadi
    HAVE_CASADI = True
except ImportError:  # noqa: D401
    casadi = None
    HAVE_CASADI = False


def require_casadi():
    if not HAVE_CASADI:
        raise ImportError("casadi is not installed: mpc_fatigue_b200.casadi_adapter needs it (pip install casadi); the "
                          "look-alike in mpc_fatigue_b200.pynocchio_casadi works without it")


def make_inverse_dynamics_callback(urdf: str, N: int, armature: float = 0.0, name: str = "gpu_inverse_dynamics"):
    """Callback tau = ID(q, qdot, qddot) for all N nodes of a horizon in one GPU launch, with the block-diagonal
    Jacobian [dtau/dq | dtau/dqdot | M] from mpcf_rnea_derivs_batch.  Inputs/outputs: vec([N, n])."""
    require_casadi()
    import torch

    from .evaluator import BatchEvaluator
    from .model import Model

    model = Model.from_urdf(urdf, armature=armature)
    ev = BatchEvaluator(model)
    n = model.n

    def to_dev(x):
        return torch.from_numpy(np.ascontiguousarray(np.array(x, dtype=np.float64).reshape(N, n).T)).cuda()

    class _Jac(casadi.Callback):
        def __init__(self):
            casadi.Callback.__init__(self)
            self.construct(name + "_jac", {})

        def get_n_in(self):
            return 4  # q, qdot, qddot, (nominal output, unused)

        def get_n_out(self):
            return 3

        def get_name_in(self, i):
            return ("q", "qdot", "qddot", "out_tau")[i]

        def get_name_out(self, i):
            return ("jac_tau_q", "jac_tau_qdot", "jac_tau_qddot")[i]

        def get_sparsity_in(self, i):
            return casadi.Sparsity.dense(N * n, 1)

        def get_sparsity_out(self, i):
            return _block_diag_sparsity(N, n)

        def eval(self, arg):
            Dq, Dv, M = ev.rnea_derivs(to_dev(arg[0]), to_dev(arg[1]), to_dev(arg[2]))
            return [_block_diag_dm(D.cpu().numpy(), N, n) for D in (Dq, Dv, M)]

    class _ID(casadi.Callback):
        def __init__(self):
            casadi.Callback.__init__(self)
            self._jac = _Jac()
            self.construct(name, {})

        def get_n_in(self):
            return 3

        def get_n_out(self):
            return 1

        def get_name_in(self, i):
            return ("q", "qdot", "qddot")[i]  # the reference's input names, src/casadi_pinocchio_bridge.hpp:78

        def get_name_out(self, i):
            return "tau"

        def get_sparsity_in(self, i):
            return casadi.Sparsity.dense(N * n, 1)

        def get_sparsity_out(self, i):
            return casadi.Sparsity.dense(N * n, 1)

        def eval(self, arg):
            tau = ev.rnea(to_dev(arg[0]), to_dev(arg[1]), to_dev(arg[2]))
            return [casadi.DM(tau.t().contiguous().cpu().numpy().reshape(-1))]

        def has_jacobian(self):
            return True

        def get_jacobian(self, jname, inames, onames, opts):
            return self._jac

    return _ID()


def _block_diag_sparsity(N: int, n: int):
    rows, cols = [], []
    for k in range(N):
        for r in range(n):
            for c in range(n):
                rows.append(k * n + r)
                cols.append(k * n + c)
    return casadi.Sparsity.triplet(N * n, N * n, rows, cols)


def _block_diag_dm(planes: np.ndarray, N: int, n: int):
    """planes [n*n, N] (row*n + col) -> block-diagonal DM."""
    out = casadi.DM(_block_diag_sparsity(N, n))
    blocks = planes.T.reshape(N, n, n)
    for k in range(N):
        out[k * n:(k + 1) * n, k * n:(k + 1) * n] = blocks[k]
    return out


def make_dyn_fatigue_step_callback(urdf: str, N: int, armature: float = 0.0, name: str = "gpu_dyn_fatigue_step", **model_kw):
    """Callback (q_next, qd_next, f_next) = dyn_fatigue_step(q, qd, tau, f, dt) for all N nodes of a horizon in one GPU launch
    (the multiple-shooting defect rows are x_{k+1} - dyn_fatigue_step(x_k, tau_k, dt)), with its Jacobian from
    mpcf_step_rk4_jvp_batch: for every (output, state/control input) pair a block-diagonal [N n, N n] matrix with the node's
    n x n block, and a dense [N n, 1] column for dt.  Vector inputs/outputs are vec([N, n]) (node-major), dt is a scalar."""
    require_casadi()
    import torch

    from .evaluator import BatchEvaluator
    from .model import Model

    model = Model.from_urdf(urdf, armature=armature, **model_kw)
    ev = BatchEvaluator(model)
    n = model.n
    in_names = ("q", "qd", "tau", "f", "dt")
    out_names = ("q_next", "qd_next", "f_next")

    def to_dev(x):
        return torch.from_numpy(np.ascontiguousarray(np.array(x, dtype=np.float64).reshape(N, n).T)).cuda()

    def vec(t):  # [n, N] device -> DM column vec([N, n])
        return casadi.DM(t.t().contiguous().cpu().numpy().reshape(-1))

    class _Jac(casadi.Callback):
        def __init__(self):
            casadi.Callback.__init__(self)
            self.construct(name + "_jac", {})

        def get_n_in(self):
            return len(in_names) + len(out_names)  # inputs, then the nominal outputs (unused)

        def get_n_out(self):
            return len(in_names) * len(out_names)  # jac_<out>_<in>, output-major

        def get_name_in(self, i):
            return in_names[i] if i < len(in_names) else "out_" + out_names[i - len(in_names)]

        def get_name_out(self, k):
            return "jac_%s_%s" % (out_names[k // len(in_names)], in_names[k % len(in_names)])

        def get_sparsity_in(self, i):
            return casadi.Sparsity.dense(1 if i == 4 else N * n, 1)

        def get_sparsity_out(self, k):
            return casadi.Sparsity.dense(N * n, 1) if k % len(in_names) == 4 else _block_diag_sparsity(N, n)

        def eval(self, arg):
            _, _, _, jac = ev.step_rk4_jvp(*(to_dev(arg[i]) for i in range(4)), float(np.array(arg[4]).reshape(-1)[0]))
            J = jac.cpu().numpy()  # [3n, 4n+1, N]
            out = []
            for o in range(3):
                for i in range(4):
                    out.append(_block_diag_dm(J[o * n:(o + 1) * n, i * n:(i + 1) * n, :].reshape(n * n, N), N, n))
                out.append(casadi.DM(np.ascontiguousarray(J[o * n:(o + 1) * n, 4 * n, :].T).reshape(-1)))
            return out

    class _Step(casadi.Callback):
        def __init__(self):
            casadi.Callback.__init__(self)
            self._jac = _Jac()
            self.construct(name, {})

        def get_n_in(self):
            return 5

        def get_n_out(self):
            return 3

        def get_name_in(self, i):
            return in_names[i]

        def get_name_out(self, i):
            return out_names[i]

        def get_sparsity_in(self, i):
            return casadi.Sparsity.dense(1 if i == 4 else N * n, 1)

        def get_sparsity_out(self, i):
            return casadi.Sparsity.dense(N * n, 1)

        def eval(self, arg):
            qn, qdn, fn = ev.step_rk4(*(to_dev(arg[i]) for i in range(4)), float(np.array(arg[4]).reshape(-1)[0]))
            return [vec(qn), vec(qdn), vec(fn)]

        def has_jacobian(self):
            return True

        def get_jacobian(self, jname, inames, onames, opts):
            return self._jac

    return _Step()


IPOPT_OPTIONS = {
    # the evaluator supplies first derivatives only (reference default: exact Hessian from the SX graph;
    # `limited-memory` appears in the reference only as a comment, python/Centauro_script/RepeatedMPCwithThermal.py:426)
    "ipopt.hessian_approximation": "limited-memory",
    "ipopt.print_level": 0,
}
