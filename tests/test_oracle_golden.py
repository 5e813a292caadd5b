"""Pin the CPU oracle against the reference's own artefacts: the stored OCP solutions
plotter/Result_{1,2,4}/solution.csv + plotter/solution.csv (committed as tests/golden/plotter_solutions.npz)
and trajectories produced by running the reference's TemperatureModel.TempSimulation
(tests/golden/thermal_known_answer.npz).  CPU only; nothing here touches /root/reference."""
import os

import numpy as np
import pytest

from mpc_fatigue_b200.model import data_urdf
from mpc_fatigue_b200.ocp import DualArmBoxOCP, f0_bound_schedule, temp_simulation
from oracle.pyoracle import Oracle
from oracle.urdf_model import load_urdf, thermal_fatigue_row

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sols():
    return np.load(os.path.join(GOLD, "plotter_solutions.npz"))


@pytest.fixture(scope="module")
def arms():
    m1, m2 = load_urdf(data_urdf("pilz6_first")), load_urdf(data_urdf("pilz6_second"))
    return (m1, Oracle(m1)), (m2, Oracle(m2))


def _T(a):
    return np.ascontiguousarray(np.asarray(a).T)


@pytest.mark.parametrize("key,N,tol", [("Result_1", 80, 1e-12), ("Result_2", 80, 1e-12), ("Result_4", 80, 5e-10), ("plotter", 50, 1e-9)])
def test_layout_and_euler_defect(sols, key, N, tol):
    s = DualArmBoxOCP.parse_solution(sols[key])
    assert s["N"] == N
    h = 2.0 / N
    # q_{k+1} = q_k + h qd_k  (Box_Pilz_6DOF2.py:463-475): IPOPT drives the defect to ~1e-13
    assert np.abs(s["q"][1:] - (s["q"][:-1] + h * s["qd"])).max() < tol
    # force equilibrium row (Box_Pilz_6DOF2.py:272-274), tolerance 1e-4
    assert np.abs(s["F_LR"][:, 2] + s["F_RR"][:, 2] - 9.81 * 30).max() < 1.1e-4


@pytest.mark.parametrize("key", ["Result_1", "Result_2", "Result_4"])
def test_fk_of_node0_hits_the_ik_targets(sols, arms, key):
    (m1, o1), (m2, o2) = arms
    s = DualArmBoxOCP.parse_solution(sols[key])
    p1, R1 = o1.fk(m1.frame_id("end_effector"), _T(s["q"][:1, :6]))
    p2, R2 = o2.fk(m2.frame_id("end_effector"), _T(s["q"][:1, 6:]))
    # targets: Box_Pilz_6DOF2.py:96-105,126,142
    assert np.abs(p1[:, 0] - [0.2, 0.6, 0.4]).max() < 2e-6
    assert np.abs(p2[:, 0] - [0.4, 0.6, 0.4]).max() < 2e-6
    assert np.abs(R1[:, 0].reshape(3, 3) - [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]).max() < 2e-3
    assert np.abs(R2[:, 0].reshape(3, 3) - [[0, 0, -1], [0, 1, 0], [1, 0, 0]]).max() < 2e-3


@pytest.mark.parametrize("key", ["Result_1", "Result_2", "Result_4", "plotter"])
def test_distance_constraint_all_nodes(sols, arms, key):
    (m1, o1), (m2, o2) = arms
    s = DualArmBoxOCP.parse_solution(sols[key])
    N = s["N"]
    p1, _ = o1.fk(m1.frame_id("end_effector"), _T(s["q"][:N, :6]))
    p2, _ = o2.fk(m2.frame_id("end_effector"), _T(s["q"][:N, 6:]))
    # |E1 - E2|^2 = L = 0.04 at every node (Box_Pilz_6DOF2.py:283-285)
    assert np.abs(((p1 - p2) ** 2).sum(0) - 0.04).max() < 1e-7


def _torques(sols, arms, key):
    (m1, o1), (m2, o2) = arms
    s = DualArmBoxOCP.parse_solution(sols[key])
    N = s["N"]
    h = 2.0 / N
    z = np.zeros((6, N))
    W1, W2 = np.vstack([s["F_LR"].T, np.zeros((3, N))]), np.vstack([s["F_RR"].T, np.zeros((3, N))])
    t1, _, _ = o1.node_eval_ref([m1.frame_id("end_effector")], -1.0, _T(s["q"][:N, :6]), _T(s["qd"][:, :6]), np.ascontiguousarray(W1), z, h)
    t2, _, _ = o2.node_eval_ref([m2.frame_id("end_effector")], -1.0, _T(s["q"][:N, 6:]), _T(s["qd"][:, 6:]), np.ascontiguousarray(W2), z, h)
    return t1, t2


def test_rnea_minus_jtw_at_active_bounds_result2(sols, arms):
    """Result_2, nodes 53..79: right-arm torques sit on the final-third bounds tau_RR0=-5, tau_RR1=+5, tau_RR2=+5."""
    _, t2 = _torques(sols, arms, "Result_2")
    dev = np.abs(t2[:3, 53:] - np.array([[-5.0], [5.0], [5.0]])).max(axis=1)
    assert dev.max() < 4e-7, dev
    assert np.abs(t2[5]).max() < 1e-12  # flange has zero inertia at the joint origin: tau_6 == 0


def test_rnea_minus_jtw_at_active_bounds_result4(sols, arms):
    """Result_4, nodes 53..79: left-arm tau_LR0 = tau_LR1 = +5 (active); right-arm rows inside their bounds."""
    t1, t2 = _torques(sols, arms, "Result_4")
    assert np.abs(t1[:2, 53:] - 5.0).max() < 2e-7
    e = 2e-7
    assert t2[0, 53:].min() > -5 - e and t2[0, 53:].max() < 5 + e
    assert t2[1, 53:].min() > -5 - e and t2[1, 53:].max() < 5 + e
    assert t2[2, 53:].min() > -10 - e and t2[2, 53:].max() < 5 + e


def test_thermal_zoh_against_reference_tempsimulation():
    """python/Libraries/TemperatureModel.py:77 run as-is (fixture) vs the oracle's ZOH map and the host twin."""
    d = np.load(os.path.join(GOLD, "thermal_known_answer.npz"))
    Ra, Rh, Rth, Tth, T, N = d["consts_Ra_Rh_Rtheta_Ttheta_T_N"]
    assert abs(Rth - 300 * 9 / 309) < 1e-12 and abs(Tth - Rth * 15) < 1e-12
    m = load_urdf(data_urdf("pilz3"), ktau=1.0)  # ktau = 1: tau plays the role of the current Ia
    assert np.allclose(m.fat[0], thermal_fatigue_row(1.0))
    orc = Oracle(m)
    h = T / N
    for k in range(4):
        Ic, Tin = d["case%d_Ic_Tin" % k]
        ref = d["case%d_Tw" % k]
        Tw = np.full((3, 1), Tin)
        got = [Tin]
        for _ in range(len(ref) - 1):
            Tw = orc.fatigue_zoh(Tw, np.full((3, 1), Ic), np.zeros((3, 1)), h)
            got.append(Tw[0, 0])
        assert np.abs(np.array(got) - ref).max() < 1e-11
        violated, tw = temp_simulation(Ic, Tin)
        assert np.abs(np.array(tw[:len(ref)]) - ref).max() < 1e-12
        assert violated == (len(ref) < int(N))


def test_f0_bound_schedule_matches_script_constants():
    # force_optimization_pilz_6DOF.py:78-89,136-148: tau0=50, alpha=2, floor=15, N=60, T=2
    b = f0_bound_schedule(60, 2.0 / 60)
    assert b[0] == 50.0 and abs(b[1] - 50 * np.exp(-2 / 30)) < 1e-14
    assert b.min() == 15.0 and np.all(np.diff(b) <= 0)
    k_floor = int(np.argmax(b == 15.0))
    assert 50 * np.exp(-2 * (k_floor - 1) / 30) > 15 >= 50 * np.exp(-2 * k_floor / 30)
