"""oracle/urdf_model.py — TEST INFRASTRUCTURE ONLY (never imported by mpc_fatigue_b200/).

URDF -> flat rigid-body model arrays for the CPU oracle.  Restates what
`pinocchio::urdf::buildModel(urdf, model, verbose)` does for the reference
(src/casadi_pinocchio_bridge.hpp:60-63, fixed base, no root joint):

* depth-first traversal from the root link, children in XML document order
  (the reference never feeds a branched URDF to the bridge, so the order is a stated choice);
* revolute / continuous / prismatic joints become 1-DOF joints; `fixed` joints merge the
  child link's inertia into the supporting moving body and only leave frames behind;
* links fixed to the world are dropped from the dynamics;
* one BODY frame per link (named like the link) and one frame per joint (named like the joint).

Every joint is normalised to "about/along local +z": a non-z axis is folded into the joint
placement (placement' = placement * R_a with R_a z = axis), which leaves tau, qdd and all frame
poses unchanged.  The C++ loader in mpc_fatigue_b200/csrc/model.cpp is an independent
implementation of the same rules; tests/test_model_loader.py compares the two.
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

# thermal constants, python/Libraries/Tmodel_library.py:9-32 (scalar ktau = 40: CentaurOCP.py:73)
RA, RH = 10.0, 2.0
RTHETA = 300.0 * 9.0 / 309.0
CTHETA = 15.0
TTHETA = RTHETA * CTHETA
KTAU = 40.0


def thermal_fatigue_row(ktau: float = KTAU) -> list[float]:
    """[lambda, kappa, ctau, cv] of  fdot = -lambda f + kappa (ctau tau^2 + cv qd^2)."""
    return [1.0 / TTHETA, RTHETA / TTHETA, RA / (ktau * ktau), 1.0 / RH]


def rpy_to_R(r: float, p: float, y: float) -> np.ndarray:
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def axis_to_R(a: np.ndarray) -> np.ndarray:
    """Rotation with R @ z = a (identity when a is already +z)."""
    a = a / np.linalg.norm(a)
    z = np.array([0.0, 0.0, 1.0])
    if np.allclose(a, z, atol=1e-14):
        return np.eye(3)
    if np.allclose(a, -z, atol=1e-14):
        return np.diag([1.0, -1.0, -1.0])
    v = np.cross(z, a)
    c = float(z @ a)
    vx = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + vx + vx @ vx / (1.0 + c)


def _floats(s: str | None, n: int, default: float = 0.0) -> np.ndarray:
    if s is None:
        return np.full(n, default)
    v = [float(x) for x in s.split()]
    assert len(v) == n, s
    return np.array(v)


@dataclass
class Model:
    n: int = 0
    parent: list[int] = field(default_factory=list)
    jtype: list[int] = field(default_factory=list)
    joint_names: list[str] = field(default_factory=list)
    Rp: list[np.ndarray] = field(default_factory=list)
    pp: list[np.ndarray] = field(default_factory=list)
    mass: list[float] = field(default_factory=list)
    mc: list[np.ndarray] = field(default_factory=list)
    Io: list[np.ndarray] = field(default_factory=list)  # 3x3 about the joint origin
    arm: list[float] = field(default_factory=list)
    fat: list[list[float]] = field(default_factory=list)
    q_lo: list[float] = field(default_factory=list)
    q_hi: list[float] = field(default_factory=list)
    v_max: list[float] = field(default_factory=list)
    tau_max: list[float] = field(default_factory=list)
    grav: tuple[float, float, float] = (0.0, 0.0, -9.81)
    frame_names: list[str] = field(default_factory=list)
    fparent: list[int] = field(default_factory=list)
    fR: list[np.ndarray] = field(default_factory=list)
    fp: list[np.ndarray] = field(default_factory=list)

    def frame_id(self, name: str) -> int:
        return self.frame_names.index(name)

    def arrays(self) -> dict[str, np.ndarray]:
        n = self.n
        Io6 = np.array([[I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]] for I in self.Io]).reshape(n, 6)
        return dict(
            parent=np.array(self.parent, dtype=np.int32),
            jtype=np.array(self.jtype, dtype=np.int32),
            Rp=np.array(self.Rp, dtype=np.float64).reshape(n, 9),
            pp=np.array(self.pp, dtype=np.float64).reshape(n, 3),
            mass=np.array(self.mass, dtype=np.float64),
            mc=np.array(self.mc, dtype=np.float64).reshape(n, 3),
            Io=Io6.astype(np.float64),
            arm=np.array(self.arm, dtype=np.float64),
            fat=np.array(self.fat, dtype=np.float64).reshape(n, 4),
            fparent=np.array(self.fparent, dtype=np.int32),
            fR=np.array(self.fR, dtype=np.float64).reshape(len(self.fparent), 9),
            fp=np.array(self.fp, dtype=np.float64).reshape(len(self.fparent), 3),
            q_lo=np.array(self.q_lo), q_hi=np.array(self.q_hi),
            v_max=np.array(self.v_max), tau_max=np.array(self.tau_max),
        )


def load_urdf(xml_text: str, armature: float = 0.0, ktau: float | list[float] = KTAU) -> Model:
    root = ET.fromstring(xml_text)
    links = {}
    for le in root.findall("link"):
        inertial = le.find("inertial")
        if inertial is None:
            links[le.get("name")] = None
            continue
        o = inertial.find("origin")
        xyz = _floats(o.get("xyz") if o is not None else None, 3)
        rpy = _floats(o.get("rpy") if o is not None else None, 3)
        m = float(inertial.find("mass").get("value"))
        ie = inertial.find("inertia")
        I = np.zeros((3, 3))
        if ie is not None:
            g = lambda k: float(ie.get(k, "0"))
            I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
        Rin = rpy_to_R(*rpy)
        links[le.get("name")] = (m, xyz, Rin @ I @ Rin.T)
    joints_of = {name: [] for name in links}
    children = set()
    for je in root.findall("joint"):
        par = je.find("parent").get("link")
        ch = je.find("child").get("link")
        joints_of[par].append(je)
        children.add(ch)
    roots = [name for name in links if name not in children]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link, found %r" % roots)

    mdl = Model()

    def add_frame(name, jidx, R, p):
        mdl.frame_names.append(name)
        mdl.fparent.append(jidx)
        mdl.fR.append(R.copy())
        mdl.fp.append(p.copy())

    def add_inertia(jidx, link_name, R, p):
        body = links[link_name]
        if body is None or jidx < 0:
            return
        m, c, Ic = body
        cj = R @ c + p
        Icj = R @ Ic @ R.T
        mdl.mass[jidx] += m
        mdl.mc[jidx] = mdl.mc[jidx] + m * cj
        mdl.Io[jidx] = mdl.Io[jidx] + Icj + m * ((cj @ cj) * np.eye(3) - np.outer(cj, cj))

    def visit(link_name, jidx, R, p):
        """(R, p): placement of `link_name`'s frame in the frame of supporting joint `jidx`."""
        add_frame(link_name, jidx, R, p)
        add_inertia(jidx, link_name, R, p)
        for je in joints_of[link_name]:
            o = je.find("origin")
            xyz = _floats(o.get("xyz") if o is not None else None, 3)
            rpy = _floats(o.get("rpy") if o is not None else None, 3)
            Rj = R @ rpy_to_R(*rpy)
            pj = p + R @ xyz
            jt = je.get("type")
            child = je.find("child").get("link")
            if jt == "fixed":
                add_frame(je.get("name"), jidx, Rj, pj)
                visit(child, jidx, Rj, pj)
                continue
            if jt not in ("revolute", "continuous", "prismatic"):
                raise ValueError("unsupported joint type %r (joint %s)" % (jt, je.get("name")))
            ax = je.find("axis")
            a = _floats(ax.get("xyz") if ax is not None and ax.get("xyz") is not None else "1 0 0", 3)
            Ra = axis_to_R(a)
            lim = je.find("limit")
            idx = mdl.n
            mdl.n += 1
            mdl.parent.append(jidx)
            mdl.jtype.append(1 if jt == "prismatic" else 0)
            mdl.joint_names.append(je.get("name"))
            mdl.Rp.append(Rj @ Ra)
            mdl.pp.append(pj.copy())
            mdl.mass.append(0.0)
            mdl.mc.append(np.zeros(3))
            mdl.Io.append(np.zeros((3, 3)))
            mdl.arm.append(float(armature))
            kt = ktau[idx] if isinstance(ktau, (list, tuple)) else ktau
            mdl.fat.append(thermal_fatigue_row(kt))
            g = (lambda k, d: float(lim.get(k, d))) if lim is not None else (lambda k, d: d)
            mdl.q_lo.append(g("lower", -math.pi))
            mdl.q_hi.append(g("upper", math.pi))
            mdl.v_max.append(g("velocity", 1.0))
            mdl.tau_max.append(g("effort", 1.0))
            # the joint's own frame keeps the URDF orientation (Pinocchio JOINT frame): Ra^T in the axis-normalised joint frame
            add_frame(je.get("name"), idx, Ra.T.copy(), np.zeros(3))
            visit(child, idx, Ra.T.copy(), np.zeros(3))

    visit(roots[0], -1, np.eye(3), np.zeros(3))
    if mdl.n == 0:
        raise ValueError("URDF has no moving joints")
    return mdl
