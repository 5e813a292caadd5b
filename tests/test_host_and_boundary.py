"""CPU-side checks of the boundary and the host logic: the C-ABI library loads and exports every symbol
include/mpcf.h declares, the C++ URDF loader agrees with the oracle's independent Python loader, error
behaviour, the CasADi-Function look-alike, bound schedules, fixture layout, synthetic batches, sharding."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from mpc_fatigue_b200 import _capi
from mpc_fatigue_b200.dist import shard_range
from mpc_fatigue_b200.model import Model, data_urdf
from mpc_fatigue_b200.ocp import DualArmBoxOCP, step_bound_table
from mpc_fatigue_b200.pynocchio_casadi import (Function, generate_forward_kin, generate_fwd_dyn_fatigue_step, generate_inv_dyn,
                                               generate_jacobian)
from mpc_fatigue_b200.synth import synth_batch
from oracle.urdf_model import load_urdf


def test_library_exports_every_symbol_of_the_header():
    hdr = open(os.path.join(ROOT, "include", "mpcf.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mpcf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(_capi.lib, name), "libmpcf.so does not export %s" % name
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)


@pytest.mark.parametrize("name", ["pilz6", "pilz3", "pilz6_first", "pilz6_second", "pilz6x2"])
def test_cpp_loader_matches_python_loader(name):
    xml = data_urdf(name)
    m, o = Model.from_urdf(xml, armature=1e-2), load_urdf(xml, armature=1e-2)
    a = o.arrays()
    assert (m.n, m.joint_names, m.frame_names) == (o.n, o.joint_names, o.frame_names)
    for k in ("parent", "jtype", "fparent", "Rp", "pp", "mass", "mc", "Io", "arm", "fat", "fR", "fp", "q_lo", "q_hi", "v_max", "tau_max"):
        assert np.abs(m.export(k).reshape(-1) - a[k].reshape(-1)).max() < 1e-15, k
    assert m.kernel_family == {"pilz6": "chain6", "pilz3": "chain3", "pilz6x2": "forest12x6"}.get(name, "chain6")


def test_fixed_joint_merge_3dof():
    m3, m6 = Model.from_urdf(data_urdf("pilz3")), Model.from_urdf(data_urdf("pilz6"))
    # joints 4-6 fixed: links 4, 5 and the flange merge into link 3
    assert m3.n == 3 and abs(m3.export("mass")[2] - m6.export("mass")[2:].sum()) < 1e-12
    assert m3.frame_id("prbt_link_5") >= 0 and m3.export("fparent")[m3.frame_id("prbt_link_5")] == 2


def test_dual_arm_tree_joint_order():
    m = Model.from_urdf(data_urdf("pilz6x2"))
    assert m.joint_names == ["prbt_joint_%d" % i for i in range(1, 7)] + ["sec_prbt_joint_%d" % i for i in range(1, 7)]
    assert m.export("parent").tolist() == [-1, 0, 1, 2, 3, 4, -1, 6, 7, 8, 9, 10]
    assert abs(m.export("pp")[6][0] - 0.7) < 1e-15  # second base offset x = +0.7


def test_error_codes_and_messages():
    with pytest.raises(ValueError, match="unsupported joint type"):
        Model.from_urdf("<robot><link name='a'/><link name='b'/><joint name='j' type='floating'><parent link='a'/><child link='b'/></joint></robot>")
    with pytest.raises(ValueError):
        Model.from_urdf("<robot><link name='a'>")
    with pytest.raises(ValueError, match="no moving joints"):
        Model.from_urdf("<robot><link name='a'/></robot>")
    with pytest.raises(ValueError):
        Model.from_urdf("not xml at all")
    m = Model.from_urdf(data_urdf("pilz6"))
    with pytest.raises(IndexError, match="unknown frame"):
        m.frame_id("no_such_body")
    with pytest.raises(ValueError):
        Model.synthetic("humanoid", 3)
    with pytest.raises(ValueError):
        Model.synthetic("chain", 65)
    # null device pointers are refused before any launch
    rc = _capi.lib.mpcf_rnea_batch(m.handle, 8, None, None, None, None, None)
    assert rc == _capi.EINVAL and "null" in _capi.last_error()


def test_synthetic_humanoid_tree():
    m = Model.synthetic("humanoid", 37, seed=7, armature=1e-2)
    par, jt = m.export("parent"), m.export("jtype")
    assert m.n == 37 and m.kernel_family == "generic64"
    assert all(par[i] < i for i in range(37))  # topological order
    assert jt[:3].tolist() == [1, 1, 1] and jt[3:].sum() == 0  # 3 prismatic + 34 revolute
    assert np.bincount(par[par >= 0], minlength=37).max() >= 4  # branched
    m2 = Model.synthetic("humanoid", 37, seed=7, armature=1e-2)
    assert np.array_equal(m.export("Rp"), m2.export("Rp")) and np.array_equal(m.export("Io"), m2.export("Io"))
    assert not np.array_equal(m.export("Rp"), Model.synthetic("humanoid", 37, seed=8).export("Rp"))


def test_synthetic_dual_arm_and_static_families():
    m = Model.synthetic("dual_arm", 14, seed=4, armature=1e-2)  # the reference's Centauro layout: two 7-DOF arms
    assert m.n == 14 and m.kernel_family == "forest14x7"
    assert m.export("parent").tolist() == [-1, 0, 1, 2, 3, 4, 5, -1, 7, 8, 9, 10, 11, 12]
    assert m.joint_names[0] == "left_joint_1" and m.joint_names[7] == "right_joint_1"
    assert m.frame_id("left_ee") >= 0 and m.frame_id("right_ee") >= 0
    assert Model.synthetic("dual_arm", 12).kernel_family == "forest12x6"
    assert Model.synthetic("dual_arm", 10).kernel_family == "generic16"
    assert Model.synthetic("chain", 7).kernel_family == "chain7"
    with pytest.raises(ValueError):
        Model.synthetic("dual_arm", 13)


def test_setters_update_model():
    m = Model.from_urdf(data_urdf("pilz6"))
    m.set_armature(0.02)
    assert np.all(m.export("arm") == 0.02)
    rows = np.tile([2.0, 0.0, 0.0, 0.0], (6, 1))  # capacity-decay instance (F0): lambda = alpha, kappa = 0
    m.set_fatigue(rows)
    assert np.array_equal(m.export("fat"), rows)


def test_function_lookalike_metadata_and_protocol():
    xml = data_urdf("pilz6_first")
    tok = generate_inv_dyn(xml)
    assert isinstance(tok, str)  # the reference returns a string that callers deserialize (force_optimization_pilz_6DOF.py:33-35)
    Idyn = Function.deserialize(tok)
    assert Idyn.name() == "inverse_dynamics" and Idyn.name_in() == ["q", "qdot", "qddot"] and Idyn.name_out() == ["tau"]
    assert Idyn.size_in("q") == (6, 1) and Idyn.size_out(0) == (6, 1) and Idyn.sparsity_out(0).is_dense()
    fk = Function.deserialize(generate_forward_kin(xml, "end_effector"))
    assert fk.name_out() == ["ee_pos", "ee_rot"] and fk.size_out("ee_rot") == (3, 3)
    J = Function.deserialize(generate_jacobian(xml, "end_effector"))
    assert J.size_out("J") == (6, 6) and J.sparsity_out("J").nnz() == 36
    step = Function.deserialize(generate_fwd_dyn_fatigue_step(xml, {"armature": 1e-2}))
    assert step.name_in() == ["q", "qd", "tau", "f", "dt"] and step.name_out() == ["q_next", "qd_next", "f_next"]
    assert step.jacobian().size_out("jac") == (18, 25)
    assert Function.deserialize(step.serialize()).name() == "dyn_fatigue_step"
    with pytest.raises(IndexError):
        Function.deserialize(generate_forward_kin(xml, "missing_body"))
    with pytest.raises(ValueError):
        Function.deserialize("garbage")
    with pytest.raises(KeyError):
        Idyn(q=[0] * 6, qdot=[0] * 6, bogus=[0] * 6)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Idyn(q=[0.0] * 6, qdot=[0.0] * 6, qddot=[0.0] * 6)


def test_casadi_adapter_degrades_cleanly_without_casadi():
    from mpc_fatigue_b200 import casadi_adapter
    if casadi_adapter.HAVE_CASADI:
        pytest.skip("casadi is installed: the adapter is exercised by the integration, not by this guard test")
    with pytest.raises(ImportError, match="casadi is not installed"):
        casadi_adapter.make_inverse_dynamics_callback(data_urdf("pilz6"), N=4)
    assert casadi_adapter.IPOPT_OPTIONS["ipopt.hessian_approximation"] == "limited-memory"


def test_step_bound_table_and_solution_layout():
    lb, ub = step_bound_table(9, [(-np.ones(2) * 50, np.ones(2) * 50), (-np.ones(2) * 20, np.ones(2) * 20), (np.array([-5, -10.0]), np.array([5, 5.0]))])
    assert lb.shape == (9, 2) and ub[0, 0] == 50 and ub[3, 0] == 20 and lb[8].tolist() == [-5, -10]
    N = 7
    v = np.arange(N * 30 + 12, dtype=float)
    s = DualArmBoxOCP.parse_solution(v)
    assert s["N"] == N and s["q"].shape == (N + 1, 12) and s["qd"][1, 0] == 30 + 12 and s["F_RR"][0, 2] == 29
    assert s["q"][N, 0] == N * 30
    with pytest.raises(ValueError):
        DualArmBoxOCP.parse_solution(v[:-1])


def test_synth_batches_are_shard_invariant_and_in_range():
    m = Model.from_urdf(data_urdf("pilz6"))
    lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
    B, N = 12, 5
    full = synth_batch(lim, 0, B, N, seed=1234, device="cpu")
    a0, a1 = shard_range(B, 0, 2)
    b0, b1 = shard_range(B, 1, 2)
    s0, s1 = synth_batch(lim, a0, a1 - a0, N, device="cpu"), synth_batch(lim, b0, b1 - b0, N, device="cpu")
    for t_full, t0, t1 in zip(full, s0, s1):
        v = t_full.reshape(6, N, B)
        assert torch.equal(v[:, :, a0:a1], t0.reshape(6, N, a1 - a0)) and torch.equal(v[:, :, b0:b1], t1.reshape(6, N, b1 - b0))
    q, qd, tau, f = full
    assert (q.amin(1) >= torch.tensor(lim["q_lo"])).all() and (q.amax(1) <= torch.tensor(lim["q_hi"])).all()
    assert qd.abs().max() <= 1.57 and (tau.abs().amax(1) <= 0.25 * torch.tensor(lim["tau_max"])).all()
    assert f.min() >= 20 and f.max() <= 80 and f.std() > 10
    assert not torch.equal(q, synth_batch(lim, 0, B, N, seed=99, device="cpu")[0])


def test_numa_binding_helpers():
    from mpc_fatigue_b200.dist import _parse_cpulist, bind_to_gpu_numa_node
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and _parse_cpulist("") == set()
    if not torch.cuda.is_available():
        before = os.sched_getaffinity(0)
        assert bind_to_gpu_numa_node(0) is None and os.sched_getaffinity(0) == before  # no topology: nothing changes


def test_shard_range_partitions():
    for B, W in ((10, 3), (8, 8), (1048576, 8), (5, 8)):
        spans = [shard_range(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)
