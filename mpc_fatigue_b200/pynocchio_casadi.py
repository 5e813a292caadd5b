"""Drop-in for the reference's pybind11 module `mpc_fatigue.pynocchio_casadi`
(bindings/python/pynocchio_casadi.cpp:11-16), which every script imports as `pin`:

    Idyn = Function.deserialize(pin.generate_inv_dyn(urdf))          # force_optimization_pilz_6DOF.py:33-35
    tau  = Idyn(q=qc_k, qdot=qcd_k, qddot=qcddot)['tau']             # :134

The three generators keep their names, arguments and the "returns a string you deserialize" protocol.  The
string is no longer a serialized CasADi SX graph but a small JSON token naming the model and the function;
`Function.deserialize` (this module's look-alike of `casadi.Function.deserialize`) turns it into a callable
with CasADi-Function call semantics that evaluates on the GPU, for one node or for any batch of nodes:

    inputs  : last axis = the CasADi vector (q[nq], qdot[nv], ...); any leading batch shape
    outputs : same leading batch shape; `ee_rot` is [..., 3, 3], `J` is [..., 6, nv] (dense, as in the reference)
    numpy / list inputs  -> numpy outputs (host path: H2D, kernel, D2H)
    torch CUDA inputs    -> torch CUDA outputs (no copies)

New (north star): `generate_fwd_dyn_fatigue_step(urdf, opts)` -> Function
`dyn_fatigue_step(q, qd, tau, f, dt) -> (q_next, qd_next, f_next)` with `.jacobian()` giving the dense
forward-mode Jacobian Function.
"""
from __future__ import annotations

import json

import numpy as np
import torch

from .evaluator import BatchEvaluator
from .model import Model

_TOKEN = "mpcf-function/1:"


def generate_inv_dyn(urdf_string: str) -> str:
    """src/casadi_pinocchio_bridge.hpp:57-85 — Function inverse_dynamics(q, qdot, qddot) -> tau."""
    return _TOKEN + json.dumps({"kind": "inverse_dynamics", "urdf": urdf_string})


def generate_forward_kin(urdf_string: str, body_name: str) -> str:
    """src/casadi_pinocchio_bridge.hpp:87-117 — Function forward_kinematics(q) -> ee_pos, ee_rot."""
    return _TOKEN + json.dumps({"kind": "forward_kinematics", "urdf": urdf_string, "frame": body_name})


def generate_jacobian(urdf_string: str, body_name: str) -> str:
    """src/casadi_pinocchio_bridge.hpp:119-153 — Function jacobian(q) -> J (6 x nv, LOCAL_WORLD_ALIGNED)."""
    return _TOKEN + json.dumps({"kind": "jacobian", "urdf": urdf_string, "frame": body_name})


def generate_fwd_dyn_fatigue_step(urdf_string: str, opts: dict | None = None) -> str:
    """North-star addition: Function dyn_fatigue_step(q, qd, tau, f, dt) -> q_next, qd_next, f_next.
    opts: armature (default 0), ktau, fatigue=(lambda, kappa, ctau, cv), gravity."""
    return _TOKEN + json.dumps({"kind": "dyn_fatigue_step", "urdf": urdf_string, "opts": opts or {}})


class Sparsity:
    """Dense sparsity pattern (the bridge builds every output with Sparsity::dense, bridge.hpp:32,45)."""

    def __init__(self, rows: int, cols: int):
        self._shape = (rows, cols)

    def is_dense(self) -> bool:
        return True

    def size1(self) -> int:
        return self._shape[0]

    def size2(self) -> int:
        return self._shape[1]

    def nnz(self) -> int:
        return self._shape[0] * self._shape[1]

    @property
    def shape(self):
        return self._shape


_model_cache: dict = {}


def _model_for(urdf: str, opts: dict) -> Model:
    key = (urdf, json.dumps(opts, sort_keys=True))
    if key not in _model_cache:
        kw = dict(opts)
        if "fatigue" in kw:
            kw["fatigue"] = tuple(kw["fatigue"])
        if "gravity" in kw:
            kw["gravity"] = tuple(kw["gravity"])
        _model_cache[key] = Model.from_urdf(urdf, **kw)
    return _model_cache[key]


class Function:
    """CasADi-Function look-alike bound to the GPU evaluator."""

    def __init__(self, spec: dict, device=None):
        self._spec = spec
        kind = spec["kind"]
        self._model = _model_for(spec["urdf"], spec.get("opts", {}))
        self._device = device
        self._ev = None
        n = self._model.n
        self._frame = None
        if kind in ("forward_kinematics", "jacobian"):
            self._frame = self._model.frame_id(spec["frame"])  # IndexError for an unknown frame, like oMf.at()
        table = {
            "inverse_dynamics": (["q", "qdot", "qddot"], [(n, 1)] * 3, ["tau"], [(n, 1)]),
            "forward_kinematics": (["q"], [(n, 1)], ["ee_pos", "ee_rot"], [(3, 1), (3, 3)]),
            "jacobian": (["q"], [(n, 1)], ["J"], [(6, n)]),
            "dyn_fatigue_step": (["q", "qd", "tau", "f", "dt"], [(n, 1)] * 4 + [(1, 1)], ["q_next", "qd_next", "f_next"], [(n, 1)] * 3),
            "jac_dyn_fatigue_step": (["q", "qd", "tau", "f", "dt"], [(n, 1)] * 4 + [(1, 1)], ["jac"], [(3 * n, 4 * n + 1)]),
        }
        if kind not in table:
            raise ValueError("unknown function kind %r" % kind)
        self._kind = kind
        self._name_in, self._size_in, self._name_out, self._size_out = table[kind]

    # ---- casadi.Function API subset ----
    @staticmethod
    def deserialize(s: str, device=None) -> "Function":
        if not isinstance(s, str) or not s.startswith(_TOKEN):
            raise ValueError("not a string produced by mpc_fatigue_b200.pynocchio_casadi.generate_*")
        return Function(json.loads(s[len(_TOKEN):]), device=device)

    def serialize(self) -> str:
        return _TOKEN + json.dumps(self._spec)

    def name(self) -> str:
        return self._kind

    def n_in(self) -> int:
        return len(self._name_in)

    def n_out(self) -> int:
        return len(self._name_out)

    def name_in(self, i: int | None = None):
        return list(self._name_in) if i is None else self._name_in[i]

    def name_out(self, i: int | None = None):
        return list(self._name_out) if i is None else self._name_out[i]

    def size_in(self, i):
        return self._size_in[self._idx(i, self._name_in)]

    def size_out(self, i):
        return self._size_out[self._idx(i, self._name_out)]

    def sparsity_in(self, i) -> Sparsity:
        return Sparsity(*self.size_in(i))

    def sparsity_out(self, i) -> Sparsity:
        return Sparsity(*self.size_out(i))

    def jacobian(self) -> "Function":
        if self._kind != "dyn_fatigue_step":
            raise NotImplementedError("forward-mode Jacobians are provided for dyn_fatigue_step")
        spec = dict(self._spec)
        spec["kind"] = "jac_dyn_fatigue_step"
        return Function(spec, device=self._device)

    @property
    def model(self) -> Model:
        return self._model

    @staticmethod
    def _idx(i, names):
        return names.index(i) if isinstance(i, str) else int(i)

    # ---- call ----
    def __call__(self, *args, **kwargs):
        if args and kwargs:
            raise TypeError("call with positional OR keyword arguments, like a casadi.Function")
        if kwargs:
            unknown = set(kwargs) - set(self._name_in)
            if unknown:
                raise KeyError("unknown input name(s) %s; inputs are %s" % (sorted(unknown), self._name_in))
            vals = [kwargs.get(nm) for nm in self._name_in]
        else:
            if len(args) != len(self._name_in):
                raise TypeError("%s takes %d inputs %s" % (self._kind, len(self._name_in), self._name_in))
            vals = list(args)
        outs = self._eval(vals)
        if kwargs:
            return dict(zip(self._name_out, outs))
        return outs[0] if len(outs) == 1 else tuple(outs)

    def _evaluator(self, dev) -> BatchEvaluator:
        if self._ev is None or self._ev.device != dev:
            self._ev = BatchEvaluator(self._model, dev)
        return self._ev

    def _eval(self, vals):
        n = self._model.n
        on_gpu = any(isinstance(v, torch.Tensor) and v.is_cuda for v in vals)
        if on_gpu:
            dev = next(v.device for v in vals if isinstance(v, torch.Tensor) and v.is_cuda)
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("mpc_fatigue_b200 evaluates on a CUDA device; none is available (no CPU fallback)")
            dev = torch.device("cuda", torch.cuda.current_device()) if self._device is None else torch.device(self._device)
        ev = self._evaluator(dev)
        # vector inputs: [..., n] (a CasADi column [n, 1] is accepted too); missing keyword inputs default to 0 like casadi
        batch = None
        cols = []
        for nm, sz, v in zip(self._name_in, self._size_in, vals):
            width = sz[0]
            if v is None:
                cols.append(None)
                continue
            t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float64))
            t = t.to(dtype=torch.float64)
            if t.dim() >= 2 and t.shape[-1] == 1 and t.shape[-2] == width and width != 1:
                t = t.squeeze(-1)  # column vector [n, 1]
            if width == 1 and (t.dim() == 0 or t.shape[-1] != 1):
                t = t.unsqueeze(-1)
            if t.shape[-1] != width:
                raise ValueError("input %r: last axis is %d, expected %d" % (nm, t.shape[-1], width))
            b = tuple(t.shape[:-1])
            if width != 1 or b not in ((), (1,)):
                if batch is None or len(b) > len(batch) or (batch in ((), (1,)) and b not in ((), (1,))):
                    batch = b
            cols.append(t)
        batch = batch if batch is not None else ()
        U = int(np.prod(batch)) if batch else 1

        def soa(t, width):
            if t is None:
                return None
            t = t.to(dev, non_blocking=True)
            t = t.expand(*batch, width) if tuple(t.shape[:-1]) != batch else t
            return t.reshape(U, width).t().contiguous()

        k = self._kind
        if k == "inverse_dynamics":
            tau = ev.rnea(soa(cols[0], n), soa(cols[1], n) if cols[1] is not None else torch.zeros((n, U), dtype=torch.float64, device=dev),
                          soa(cols[2], n))
            outs = [tau.t().reshape(*batch, n)]
        elif k == "forward_kinematics":
            pos, rot = ev.fk(self._frame, soa(cols[0], n))
            outs = [pos.t().reshape(*batch, 3), rot.t().reshape(*batch, 3, 3)]
        elif k == "jacobian":
            J = ev.jacobian(self._frame, soa(cols[0], n))
            outs = [J.t().reshape(*batch, 6, n)]
        else:
            dt = cols[4]
            if dt is None:
                raise ValueError("dt is required")
            if dt.numel() == 1:
                dt_arg = float(dt.reshape(-1)[0].item())
            else:
                dt_arg = dt.to(dev).expand(*batch, 1).reshape(U).contiguous()
            a = [soa(cols[i], n) if cols[i] is not None else torch.zeros((n, U), dtype=torch.float64, device=dev) for i in range(4)]
            if k == "dyn_fatigue_step":
                qn, qdn, fn = ev.step_rk4(*a, dt_arg)
                outs = [x.t().reshape(*batch, n) for x in (qn, qdn, fn)]
            else:
                _, _, _, jac = ev.step_rk4_jvp(*a, dt_arg)
                outs = [jac.permute(2, 0, 1).reshape(*batch, 3 * n, 4 * n + 1)]
        if not on_gpu:
            outs = [o.cpu().numpy() for o in outs]
        return outs
