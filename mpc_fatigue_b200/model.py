"""Model handle: URDF / synthetic tree -> immutable native model (mpcf_model).

Replaces the `urdf::parseURDF` + `pinocchio::urdf::buildModel` prologue that every generator of the
reference bridge repeats (src/casadi_pinocchio_bridge.hpp:60-63, 91-94, 123-126).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "models")


def data_urdf(name: str) -> str:
    """Text of a bundled minimal URDF (kinematic + inertial data only), e.g. 'pilz6'."""
    with open(os.path.join(_DATA, name + ".urdf")) as fh:
        return fh.read()


def make_opts(armature: float = 0.0, gravity=(0.0, 0.0, -9.81), ktau: float | None = None,
              fatigue: tuple[float, float, float, float] | None = None) -> _capi.Opts:
    o = _capi.Opts()
    _capi.lib.mpcf_opts_default(C.byref(o))
    o.armature = float(armature)
    o.gravity = (C.c_double * 3)(*[float(g) for g in gravity])
    if ktau is not None:
        o.ctau = 10.0 / (float(ktau) ** 2)  # Ra / ktau^2, python/Libraries/Tmodel_library.py:9,32
    if fatigue is not None:
        o.lambda_, o.kappa, o.ctau, o.cv = [float(v) for v in fatigue]
    return o


class Model:
    _INT_FIELDS = ("parent", "jtype", "fparent")
    _SHAPES = {"Rp": 9, "pp": 3, "mass": 1, "mc": 3, "Io": 6, "arm": 1, "fat": 4, "fR": 9, "fp": 3,
               "q_lo": 1, "q_hi": 1, "v_max": 1, "tau_max": 1}

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        nq, nv, nb, nf = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _capi.check(_capi.lib.mpcf_model_info(self._h, C.byref(nq), C.byref(nv), C.byref(nb), C.byref(nf)))
        self.nq, self.nv, self.nframes = nq.value, nv.value, nf.value
        self.n = self.nv
        self.joint_names = [_capi.lib.mpcf_joint_name(self._h, i).decode() for i in range(self.n)]
        self.frame_names = [_capi.lib.mpcf_frame_name(self._h, i).decode() for i in range(self.nframes)]

    # ---- constructors ----
    @classmethod
    def from_urdf(cls, xml: str, armature: float = 0.0, **kw) -> "Model":
        data = xml.encode() if isinstance(xml, str) else bytes(xml)
        opts = make_opts(armature=armature, **kw)
        out = C.c_void_p()
        _capi.check(_capi.lib.mpcf_model_create_from_urdf(data, len(data), C.byref(opts), C.byref(out)))
        return cls(out.value)

    @classmethod
    def synthetic(cls, kind: str, ndof: int, seed: int = 1, armature: float = 0.0, **kw) -> "Model":
        kinds = {"chain": _capi.SYNTH_CHAIN, "humanoid": _capi.SYNTH_HUMANOID, "dual_arm": _capi.SYNTH_DUAL_ARM}
        if kind not in kinds:
            raise ValueError("unknown synthetic kind %r" % kind)
        opts = make_opts(armature=armature, **kw)
        out = C.c_void_p()
        _capi.check(_capi.lib.mpcf_model_create_synthetic(kinds[kind], int(ndof), int(seed), C.byref(opts), C.byref(out)))
        return cls(out.value)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _capi is not None and getattr(_capi, "lib", None) is not None:  # module globals vanish at interpreter exit
            _capi.lib.mpcf_model_destroy(h)

    # ---- queries ----
    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def kernel_family(self) -> str:
        return _capi.lib.mpcf_model_kernel_family(self._h).decode()

    def frame_id(self, name: str) -> int:
        rc = _capi.lib.mpcf_frame_id(self._h, name.encode())
        _capi.check(rc)
        return rc

    def export(self, field: str) -> np.ndarray:
        if field in self._INT_FIELDS:
            cnt = self.nframes if field == "fparent" else self.n
            buf = np.empty(cnt, dtype=np.int32)
        elif field == "grav":
            buf = np.empty(3, dtype=np.float64)
        else:
            rows = self.nframes if field in ("fR", "fp") else self.n
            w = self._SHAPES[field]
            buf = np.empty((rows, w) if w > 1 else rows, dtype=np.float64)
        rc = _capi.lib.mpcf_model_export(self._h, field.encode(), buf.ctypes.data_as(C.c_void_p), buf.nbytes)
        _capi.check(rc)
        assert rc == buf.nbytes, (field, rc, buf.nbytes)
        return buf

    def set_armature(self, arm) -> None:
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(arm, dtype=np.float64), (self.n,)))
        _capi.check(_capi.lib.mpcf_model_set_armature(self._h, a.ctypes.data_as(C.POINTER(C.c_double))))

    def set_fatigue(self, rows) -> None:
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(rows, dtype=np.float64), (self.n, 4)))
        _capi.check(_capi.lib.mpcf_model_set_fatigue(self._h, a.ctypes.data_as(C.POINTER(C.c_double))))

    def set_coupling(self, coupling) -> None:
        """Couple the two arms' fatigue through a shared box load: `coupling` = (ee_frame_arm0, ee_frame_arm1, weight) with frame
        indices (or names) and the box weight m g; None removes it (include/mpcf.h: mpcf_model_set_coupling)."""
        if coupling is None:
            _capi.check(_capi.lib.mpcf_model_set_coupling(self._h, None))
            self.coupling = None
            return
        f0, f1, w = coupling
        fr = [self.frame_id(f) if isinstance(f, str) else int(f) for f in (f0, f1)]
        c = _capi.Coupling((C.c_int * 2)(*fr), float(w))
        _capi.check(_capi.lib.mpcf_model_set_coupling(self._h, C.byref(c)))
        self.coupling = (fr[0], fr[1], float(w))
