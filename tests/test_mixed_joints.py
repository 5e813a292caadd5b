"""A hand-written URDF (tests/golden/mixed_joints.urdf — not from the reference) with what the Pilz models do not have:
prismatic and continuous joints, joint axes off the local z, a fixed joint whose child carries inertia, a branch.

CPU: the C++ and the Python loader agree, and the oracle's gravity torques equal the gradient of the potential energy
computed independently from frame kinematics and the URDF's own inertial origins (this pins axis normalisation, the
fixed-joint merge and the sign conventions without using the dynamics code).  GPU: parity on the same model."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from conftest import ROOT, oracle_model_from_export, random_inputs, rel_err_rows
from mpc_fatigue_b200.model import Model
from oracle.pyoracle import Oracle
from oracle.urdf_model import load_urdf

XML = open(os.path.join(ROOT, "tests", "golden", "mixed_joints.urdf")).read()


def test_loaders_agree_on_mixed_joint_types():
    m, o = Model.from_urdf(XML, armature=1e-3), load_urdf(XML, armature=1e-3)
    a = o.arrays()
    assert m.n == 3 and m.joint_names == ["slide", "swing", "spin"] and m.kernel_family == "generic16"
    assert m.export("jtype").tolist() == [1, 0, 0] and m.export("parent").tolist() == [-1, 0, 0]
    for k in ("parent", "jtype", "fparent", "Rp", "pp", "mass", "mc", "Io", "fR", "fp", "q_lo", "q_hi", "v_max", "tau_max"):
        assert np.abs(m.export(k).reshape(-1).astype(float) - a[k].reshape(-1).astype(float)).max() < 1e-15, k
    assert abs(m.export("mass")[1] - 1.9) < 1e-15  # arm + the tool behind the fixed joint


def _potential_energy(orc, om, q):
    """U(q) = -sum_i m_i g . c_i(q) with c_i the world COM of every URDF link that has mass, from frame FK only."""
    g = np.array([0.0, 0.0, -9.81])
    U = np.zeros(q.shape[1])
    for link in ET.fromstring(XML).findall("link"):
        ine = link.find("inertial")
        if ine is None:
            continue
        mass = float(ine.find("mass").get("value"))
        org = ine.find("origin")
        com = np.array([float(v) for v in (org.get("xyz") if org is not None and org.get("xyz") else "0 0 0").split()])
        pos, rot = orc.fk(om.frame_names.index(link.get("name")), q)
        c = pos + np.einsum("iju,j->iu", rot.reshape(3, 3, -1), com)
        U -= mass * (g @ c)
    return U


def test_gravity_torque_is_the_potential_gradient():
    om = load_urdf(XML, armature=0.0)
    orc = Oracle(om)
    rng = np.random.default_rng(2)
    q = np.ascontiguousarray(rng.uniform(-1.0, 1.0, (3, 40)))
    tau_g = orc.rnea(q, np.zeros_like(q), None)
    h = 1e-6
    for j in range(3):
        dq = np.zeros_like(q)
        dq[j] = h
        dU = (_potential_energy(orc, om, q + dq) - _potential_energy(orc, om, q - dq)) / (2 * h)
        assert np.abs(tau_g[j] - dU).max() < 1e-7 * max(1.0, np.abs(dU).max()), j


@pytest.mark.gpu
def test_gpu_parity_on_mixed_joint_model():
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    m = Model.from_urdf(XML, armature=1e-3)
    om = oracle_model_from_export(m)
    orc, ev = Oracle(om), BatchEvaluator(m)
    U = 97
    q, qd, tau, f, qdd = random_inputs(om, U, seed=5)
    d = [torch.from_numpy(a).cuda() for a in (q, qd, tau, f, qdd)]
    assert rel_err_rows(ev.rnea(d[0], d[1], d[4]).cpu().numpy(), orc.rnea(q, qd, qdd)) < 1e-9
    assert rel_err_rows(ev.aba(d[0], d[1], d[2]).cpu().numpy(), orc.aba(q, qd, tau)) < 1e-9
    for fr in range(m.nframes):
        pos, rot = orc.fk(fr, q)
        gp, gr = ev.fk(fr, d[0])
        assert np.abs(gp.cpu().numpy() - pos).max() < 1e-12 and np.abs(gr.cpu().numpy() - rot).max() < 1e-12
        assert np.abs(ev.jacobian(fr, d[0]).cpu().numpy() - orc.jacobian(fr, q)).max() < 1e-12
    ref = orc.step_rk4_jvp(q, qd, tau, f, 0.01)
    got = ev.step_rk4_jvp(d[0], d[1], d[2], d[3], 0.01)
    for g_, r_ in zip(got, ref):
        assert np.abs(g_.cpu().numpy() - r_).max() < 1e-9 * max(1.0, np.abs(r_).max())


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _energies(orc, om, xml, q, qd):
    """(2 x kinetic energy, potential energy) of every unit from the URDF text and frame kinematics only."""
    n, U = q.shape
    g = np.array([0.0, 0.0, -9.81])
    ke2, pe = np.zeros(U), np.zeros(U)
    for link in ET.fromstring(xml).findall("link"):
        ine = link.find("inertial")
        if ine is None or link.get("name") not in om.frame_names:
            continue
        mass = float(ine.find("mass").get("value"))
        org = ine.find("origin")
        xyz = np.array([float(v) for v in ((org.get("xyz") if org is not None else None) or "0 0 0").split()])
        rpy = [float(v) for v in ((org.get("rpy") if org is not None else None) or "0 0 0").split()]
        I = ine.find("inertia")
        Ic = np.array([[float(I.get("ixx")), float(I.get("ixy")), float(I.get("ixz"))],
                       [float(I.get("ixy")), float(I.get("iyy")), float(I.get("iyz"))],
                       [float(I.get("ixz")), float(I.get("iyz")), float(I.get("izz"))]])
        Ic = _rpy(*rpy) @ Ic @ _rpy(*rpy).T  # inertia in link axes
        fr = om.frame_names.index(link.get("name"))
        pos, rot = orc.fk(fr, q)
        J = orc.jacobian(fr, q).reshape(6, n, U)  # LOCAL_WORLD_ALIGNED at the link origin
        tw = np.einsum("rju,ju->ru", J, qd)
        R = rot.reshape(3, 3, U)
        w = tw[3:]
        r_c = np.einsum("iju,j->ui", R, xyz)
        vc = tw[:3] + np.cross(w.T, r_c).T
        wl = np.einsum("jiu,ju->iu", R, w)  # angular velocity in link axes
        ke2 += mass * (vc * vc).sum(0) + np.einsum("iu,ij,ju->u", wl, Ic, wl)
        pe -= mass * (g @ (pos + r_c.T))
    return ke2, pe


@pytest.mark.parametrize("which", ["mixed", "pilz6"])
def test_mass_matrix_and_power_balance_from_frame_kinematics(which):
    """Two checks of the oracle's inverse dynamics against quantities built from the URDF text and frame kinematics only:
    q̇ᵀ M(q) q̇ = sum over links of m |v_c|² + ωᵀ I_c ω (pins the inertial terms), and along any motion
    q̇ᵀ ID(q, q̇, q̈) = d/dt (kinetic + potential energy) (pins the Coriolis / centrifugal and gravity terms)."""
    from mpc_fatigue_b200.model import data_urdf
    xml = XML if which == "mixed" else data_urdf("pilz6")
    om = load_urdf(xml, armature=0.0)
    orc = Oracle(om)
    n, U = om.n, 12
    rng = np.random.default_rng(9)
    q = np.ascontiguousarray(rng.uniform(-1.0, 1.0, (n, U)))
    qd = np.ascontiguousarray(rng.uniform(-1.0, 1.0, (n, U)))
    qdd = np.ascontiguousarray(rng.uniform(-2.0, 2.0, (n, U)))
    z = np.zeros_like(q)
    bias = orc.rnea(q, z, None)
    M = np.stack([orc.rnea(q, z, np.ascontiguousarray(np.eye(n)[:, [j]].repeat(U, 1))) - bias for j in range(n)], axis=1)  # [n, n, U]
    ke2, _ = _energies(orc, om, xml, q, qd)
    assert np.abs(np.einsum("iu,iju,ju->u", qd, M, qd) - ke2).max() < 1e-10 * np.abs(ke2).max()
    h = 1e-5
    E = []
    for sgn in (+1.0, -1.0):
        qs = np.ascontiguousarray(q + sgn * h * qd + 0.5 * h * h * qdd)
        k2, pe = _energies(orc, om, xml, qs, np.ascontiguousarray(qd + sgn * h * qdd))
        E.append(0.5 * k2 + pe)
    power = (qd * orc.rnea(q, qd, qdd)).sum(0)
    dE = (E[0] - E[1]) / (2 * h)
    assert np.abs(power - dE).max() < 1e-6 * max(1.0, np.abs(dE).max())


@pytest.mark.gpu
def test_gpu_power_balance_at_scale():
    """The same power balance with every term from the CUDA path (RNEA, frame poses and Jacobians through the C-ABI) on 2^20
    units of the Pilz 6-DOF model: a size-independent property that involves neither the oracle nor any CPU arithmetic."""
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import data_urdf
    xml = data_urdf("pilz6")
    m = Model.from_urdf(xml, armature=0.0)
    ev = BatchEvaluator(m)
    n, U = m.n, 1 << 20
    gen = torch.Generator(device="cuda").manual_seed(3)
    rnd = lambda s: (torch.rand((n, U), dtype=torch.float64, device="cuda", generator=gen) * 2.0 - 1.0) * s
    q, qd, qdd = rnd(1.5), rnd(1.0), rnd(2.0)
    links = []
    for link in ET.fromstring(xml).findall("link"):
        ine = link.find("inertial")
        if ine is None or link.get("name") not in m.frame_names:
            continue
        org, I = ine.find("origin"), ine.find("inertia")
        xyz = [float(v) for v in ((org.get("xyz") if org is not None else None) or "0 0 0").split()]
        rpy = [float(v) for v in ((org.get("rpy") if org is not None else None) or "0 0 0").split()]
        Ic = np.array([[float(I.get("ixx")), float(I.get("ixy")), float(I.get("ixz"))],
                       [float(I.get("ixy")), float(I.get("iyy")), float(I.get("iyz"))],
                       [float(I.get("ixz")), float(I.get("iyz")), float(I.get("izz"))]])
        links.append((m.frame_id(link.get("name")), float(ine.find("mass").get("value")), torch.tensor(xyz, dtype=torch.float64, device="cuda"),
                      torch.from_numpy(_rpy(*rpy) @ Ic @ _rpy(*rpy).T).cuda()))

    def energy(qs, qds):
        E = torch.zeros(U, dtype=torch.float64, device="cuda")
        for fr, mass, xyz, Ic in links:
            pos, rot = ev.fk(fr, qs)
            tw = torch.einsum("rju,ju->ru", ev.jacobian(fr, qs).view(6, n, U), qds)
            R = rot.view(3, 3, U)
            w = tw[3:]
            r_c = torch.einsum("iju,j->iu", R, xyz)
            vc = tw[:3] + torch.linalg.cross(w, r_c, dim=0)
            wl = torch.einsum("jiu,ju->iu", R, w)
            E += 0.5 * (mass * (vc * vc).sum(0) + torch.einsum("iu,ij,ju->u", wl, Ic, wl)) + mass * 9.81 * (pos[2] + r_c[2])
        return E

    h = 1e-5
    Ep = energy((q + h * qd + 0.5 * h * h * qdd).contiguous(), (qd + h * qdd).contiguous())
    Em = energy((q - h * qd + 0.5 * h * h * qdd).contiguous(), (qd - h * qdd).contiguous())
    dE = (Ep - Em) / (2 * h)
    power = (qd * ev.rnea(q, qd, qdd)).sum(0)
    assert float((power - dE).abs().max()) < 1e-6 * max(1.0, float(dE.abs().max()))


@pytest.mark.gpu
def test_gpu_kinetic_energy_identity_at_scale():
    """qd^T M(q) qd = sum over links of m |v_c|^2 + w^T I_c w with M from the CUDA inverse-dynamics derivative kernel
    (mpcf_rnea_derivs_batch) and the right-hand side from the CUDA frame poses / Jacobians and the URDF text: pins the inertial
    (M qdd) terms of the GPU path, which no reference fixture reaches (every reference call has qdd = 0), on 2^18 units."""
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import data_urdf
    xml = data_urdf("pilz6")
    m = Model.from_urdf(xml, armature=0.0)
    ev = BatchEvaluator(m)
    n, U = m.n, 1 << 18
    gen = torch.Generator(device="cuda").manual_seed(4)
    rnd = lambda s: (torch.rand((n, U), dtype=torch.float64, device="cuda", generator=gen) * 2.0 - 1.0) * s
    q, qd = rnd(1.5), rnd(1.0)
    _, _, M = ev.rnea_derivs(q, qd)
    lhs = torch.einsum("iu,iju,ju->u", qd, M.view(n, n, U), qd)
    rhs = torch.zeros(U, dtype=torch.float64, device="cuda")
    for link in ET.fromstring(xml).findall("link"):
        ine = link.find("inertial")
        if ine is None or link.get("name") not in m.frame_names:
            continue
        org, I = ine.find("origin"), ine.find("inertia")
        xyz = torch.tensor([float(v) for v in ((org.get("xyz") if org is not None else None) or "0 0 0").split()], dtype=torch.float64, device="cuda")
        rpy = [float(v) for v in ((org.get("rpy") if org is not None else None) or "0 0 0").split()]
        Ic = np.array([[float(I.get("ixx")), float(I.get("ixy")), float(I.get("ixz"))],
                       [float(I.get("ixy")), float(I.get("iyy")), float(I.get("iyz"))],
                       [float(I.get("ixz")), float(I.get("iyz")), float(I.get("izz"))]])
        Ic = torch.from_numpy(_rpy(*rpy) @ Ic @ _rpy(*rpy).T).cuda()
        fr = m.frame_id(link.get("name"))
        if m.export("fparent")[fr] < 0:
            continue  # fixed to the world: no kinetic energy
        mass = float(ine.find("mass").get("value"))
        _, rot = ev.fk(fr, q)
        tw = torch.einsum("rju,ju->ru", ev.jacobian(fr, q).view(6, n, U), qd)
        R = rot.view(3, 3, U)
        w = tw[3:]
        vc = tw[:3] + torch.linalg.cross(w, torch.einsum("iju,j->iu", R, xyz), dim=0)
        wl = torch.einsum("jiu,ju->iu", R, w)
        rhs += mass * (vc * vc).sum(0) + torch.einsum("iu,ij,ju->u", wl, Ic, wl)
    assert float((lhs - rhs).abs().max()) < 1e-10 * float(rhs.abs().max())


# ---- joint frames keep the URDF orientation (model.cpp / urdf_model.py registered them as identity) ----
def _rodrigues(axis, ang):
    a = np.asarray(axis, dtype=float) / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def _rpy(r, p, y):
    return _rodrigues([0, 0, 1], y) @ _rodrigues([0, 1, 0], p) @ _rodrigues([1, 0, 0], r)


def _hand_joint_frames(q):
    """World pose of the JOINT frames `slide`, `swing`, `spin` of mixed_joints.urdf written out by hand from the URDF text
    (Pinocchio semantics: joint frame = parent placement * <origin> * motion about the URDF <axis>; no axis normalisation)."""
    R0, p0 = _rpy(0, 0.3, 0), np.array([0.0, 0.0, 0.2])
    p_slide = p0 + R0 @ (q[0] * np.array([1.0, 0.0, 0.0]))       # prismatic along x: orientation unchanged
    R_swing = R0 @ _rpy(0.2, 0, -0.4) @ _rodrigues([0, 1, 0], q[1])
    p_swing = p_slide + R0 @ np.array([0.1, 0.0, 0.1])
    R_spin = R0 @ _rpy(1.2, 0, 0) @ _rodrigues([0.6, 0, 0.8], q[2])
    p_spin = p_slide + R0 @ np.array([-0.1, 0.05, 0.0])
    return {"slide": (R0, p_slide), "swing": (R_swing, p_swing), "spin": (R_spin, p_spin)}


def test_joint_frames_keep_the_urdf_orientation():
    om = load_urdf(XML, armature=0.0)
    orc = Oracle(om)
    m = Model.from_urdf(XML)
    # both loaders register the joint frames with Ra^T, and agree
    assert np.abs(m.export("fR").reshape(-1) - om.arrays()["fR"].reshape(-1)).max() < 1e-15
    rng = np.random.default_rng(11)
    q = np.ascontiguousarray(rng.uniform(-1.0, 1.0, (3, 7)))
    for name in ("slide", "swing", "spin"):
        # the joint frame is the second frame of that name? no: link frames carry link names; joint names are unique here
        fid = om.frame_names.index(name)
        pos, rot = orc.fk(fid, q)
        for u in range(q.shape[1]):
            R, p = _hand_joint_frames(q[:, u])[name]
            assert np.abs(rot[:, u].reshape(3, 3) - R).max() < 1e-13, name
            assert np.abs(pos[:, u] - p).max() < 1e-13, name


@pytest.mark.gpu
def test_gpu_joint_frames_keep_the_urdf_orientation():
    import torch
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    m = Model.from_urdf(XML)
    ev = BatchEvaluator(m)
    rng = np.random.default_rng(12)
    q = np.ascontiguousarray(rng.uniform(-1.0, 1.0, (3, 33)))
    dq = torch.from_numpy(q).cuda()
    for name in ("slide", "swing", "spin"):
        pos, rot = ev.fk(m.frame_id(name), dq)
        pos, rot = pos.cpu().numpy(), rot.cpu().numpy()
        for u in range(q.shape[1]):
            R, p = _hand_joint_frames(q[:, u])[name]
            assert np.abs(rot[:, u].reshape(3, 3) - R).max() < 1e-13 and np.abs(pos[:, u] - p).max() < 1e-13, name
