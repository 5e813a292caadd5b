"""Test infrastructure (like everything under oracle/): builds the oracle's model from the arrays a product model exports, and
seeded inputs in the value ranges of SURVEY.md §8(d).  Used by tests/ and by __graft_entry__.smoke()."""
import numpy as np


def oracle_model_from_export(m):
    """oracle.urdf_model.Model built from the arrays the product model exports (synthetic trees)."""
    from oracle.urdf_model import Model as OModel
    om = OModel()
    om.n = m.n
    om.parent = m.export("parent").tolist()
    om.jtype = m.export("jtype").tolist()
    om.joint_names = list(m.joint_names)
    om.Rp = list(m.export("Rp").reshape(m.n, 3, 3))
    om.pp = list(m.export("pp").reshape(m.n, 3))
    om.mass = m.export("mass").reshape(m.n).tolist()
    om.mc = list(m.export("mc").reshape(m.n, 3))
    Io = m.export("Io").reshape(m.n, 6)
    om.Io = [np.array([[a[0], a[1], a[2]], [a[1], a[3], a[4]], [a[2], a[4], a[5]]]) for a in Io]
    om.arm = m.export("arm").reshape(m.n).tolist()
    om.fat = m.export("fat").reshape(m.n, 4).tolist()
    om.q_lo = m.export("q_lo").reshape(m.n).tolist()
    om.q_hi = m.export("q_hi").reshape(m.n).tolist()
    om.v_max = m.export("v_max").reshape(m.n).tolist()
    om.tau_max = m.export("tau_max").reshape(m.n).tolist()
    om.grav = tuple(m.export("grav").tolist())
    om.frame_names = list(m.frame_names)
    om.fparent = m.export("fparent").tolist()
    om.fR = list(m.export("fR").reshape(m.nframes, 3, 3))
    om.fp = list(m.export("fp").reshape(m.nframes, 3))
    return om


def random_inputs(om, U, seed=0):
    """Seeded inputs in the value ranges of SURVEY.md §8(d); arrays are [n, U] float64 C-contiguous."""
    rng = np.random.default_rng(seed)
    n = om.n
    lo, hi = np.array(om.q_lo)[:, None], np.array(om.q_hi)[:, None]
    q = rng.uniform(lo, hi, (n, U))
    qd = rng.uniform(-1.0, 1.0, (n, U)) * np.array(om.v_max)[:, None]
    tau = rng.uniform(-1.0, 1.0, (n, U)) * np.array(om.tau_max)[:, None] * 0.25
    f = rng.uniform(20.0, 80.0, (n, U))
    qdd = rng.uniform(-3.0, 3.0, (n, U))
    return [np.ascontiguousarray(a) for a in (q, qd, tau, f, qdd)]
