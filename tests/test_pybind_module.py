"""`import mpc_fatigue.pynocchio_casadi as pin` — the reference's own import line (python/Libraries/Centauro_functions.py:2)
resolves to the compiled pybind11 module built from bindings/python/pynocchio_casadi.cpp against libmpcf.so."""
import json

import numpy as np
import pytest

from conftest import random_inputs, rel_err_rows
from mpc_fatigue_b200.model import data_urdf

XML = data_urdf("pilz6_first")


def test_reference_import_line_and_generators():
    import mpc_fatigue.pynocchio_casadi as pin
    import mpc_fatigue_b200.pynocchio_casadi as pypin
    assert pin.__file__.endswith(".so")  # compiled module, not a Python alias
    # the three generators of bindings/python/pynocchio_casadi.cpp:14-16, same signatures, str in -> str out
    for native, py_, args in ((pin.generate_inv_dyn, pypin.generate_inv_dyn, (XML,)),
                              (pin.generate_forward_kin, pypin.generate_forward_kin, (XML, "end_effector")),
                              (pin.generate_jacobian, pypin.generate_jacobian, (XML, "end_effector"))):
        s = native(*args)
        assert isinstance(s, str) and s.startswith("mpcf-function/1:")
        assert json.loads(s.split(":", 1)[1]) == json.loads(py_(*args).split(":", 1)[1])
    f = pin.Function.deserialize(pin.generate_inv_dyn(XML))
    assert f.name_in() == ["q", "qdot", "qddot"] and f.name_out() == ["tau"] and f.size_out("tau") == (6, 1)
    j = pin.Function.deserialize(pin.generate_jacobian(XML, "end_effector"))
    assert j.size_out("J") == (6, 6) and j.sparsity_out(0).is_dense()


def test_generators_fail_loudly_at_generation_time():
    import mpc_fatigue.pynocchio_casadi as pin
    with pytest.raises(IndexError):  # the reference: oMf.at(nframes) throws (bridge.hpp:103-107)
        pin.generate_forward_kin(XML, "no_such_frame")
    with pytest.raises(ValueError):  # the reference: parseURDF returns null, unchecked (bridge.hpp:60-63)
        pin.generate_inv_dyn("<robot name='x'><link name='a'/>")
    with pytest.raises(ValueError):
        pin.generate_jacobian(XML.replace('type="revolute"', 'type="planar"', 1), "end_effector")


def test_native_model_handle():
    import mpc_fatigue.pynocchio_casadi as pin
    m = pin.Model(XML, 1e-2)
    assert m.nv == 6 and m.kernel_family() == "chain6" and m.joint_names()[0] == "prbt_joint_1"
    assert m.frame_id("end_effector") >= 0
    with pytest.raises(IndexError):
        m.frame_id("nope")
    with pytest.raises(ValueError):  # null array argument -> MPCF_EINVAL -> ValueError, nothing launched
        m.rnea(4, 0, 0, 0, 0, 0)


@pytest.mark.gpu
def test_native_model_batch_calls_match_the_oracle():
    import torch
    import mpc_fatigue.pynocchio_casadi as pin
    from oracle.pyoracle import Oracle
    from oracle.urdf_model import load_urdf
    om = load_urdf(XML, armature=1e-2)
    orc = Oracle(om)
    m = pin.Model(XML, 1e-2)
    U = 193
    q, qd, tau, f, qdd = random_inputs(om, U, seed=21)
    d = {k: torch.from_numpy(v).cuda() for k, v in dict(q=q, qd=qd, tau=tau, f=f, qdd=qdd).items()}
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty((6, U), dtype=torch.float64, device="cuda")
    m.rnea(U, d["q"].data_ptr(), d["qd"].data_ptr(), d["qdd"].data_ptr(), out.data_ptr(), st)
    assert rel_err_rows(out.cpu().numpy(), orc.rnea(q, qd, qdd)) < 1e-9
    fr = m.frame_id("end_effector")
    pos, rot, J = (torch.empty((k, U), dtype=torch.float64, device="cuda") for k in (3, 9, 36))
    m.fk(fr, U, d["q"].data_ptr(), pos.data_ptr(), rot.data_ptr(), st)
    m.jacobian(fr, U, d["q"].data_ptr(), J.data_ptr(), st)
    rp, rr = orc.fk(fr, q)
    assert np.abs(pos.cpu().numpy() - rp).max() < 1e-12 and np.abs(rot.cpu().numpy() - rr).max() < 1e-12
    assert np.abs(J.cpu().numpy() - orc.jacobian(fr, q)).max() < 1e-12
    qn, qdn, fn = (torch.empty((6, U), dtype=torch.float64, device="cuda") for _ in range(3))
    jac = torch.empty((18, 25, U), dtype=torch.float64, device="cuda")
    m.step_rk4_jvp(U, d["q"].data_ptr(), d["qd"].data_ptr(), d["tau"].data_ptr(), d["f"].data_ptr(), 0.02, qn.data_ptr(), qdn.data_ptr(),
                   fn.data_ptr(), jac.data_ptr(), st)
    rq, rqd, rf, rj = orc.step_rk4_jvp(q, qd, tau, f, 0.02)
    assert rel_err_rows(qn.cpu().numpy(), rq) < 1e-9 and rel_err_rows(qdn.cpu().numpy(), rqd) < 1e-9 and rel_err_rows(fn.cpu().numpy(), rf) < 1e-9
    assert rel_err_rows(jac.cpu().numpy().reshape(450, U), rj.reshape(450, U)) < 1e-9
    # the Function look-alike reached through the reference's import line evaluates on the GPU too
    Idyn = pin.Function.deserialize(pin.generate_inv_dyn(XML))
    tau_ref = Idyn(q=q.T, qdot=qd.T, qddot=qdd.T)["tau"]
    assert rel_err_rows(np.asarray(tau_ref).T, Oracle(load_urdf(XML)).rnea(q, qd, qdd)) < 1e-9
