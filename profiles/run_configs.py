#!/usr/bin/env python3
"""Timing of the non-headline configs (C3 dual-arm 12-DOF, C4 synthetic 37-DOF tree) at reduced batch sizes.
    python profiles/run_configs.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mpc_fatigue_b200.evaluator import BatchEvaluator
from mpc_fatigue_b200.model import Model, data_urdf
from mpc_fatigue_b200.synth import synth_batch


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, m, B, N, dt in (("C3 pilz6x2 (forest12x6)", Model.from_urdf(data_urdf("pilz6x2"), armature=1e-2), 8192, 100, 0.02),
                          ("Centauro-like dual arm 2x7 (forest14x7)", Model.synthetic("dual_arm", 14, seed=3, armature=1e-2), 8192, 40, 0.5 / 40),
                          ("C4 humanoid37 (generic64)", Model.synthetic("humanoid", 37, seed=7, armature=1e-2), 8192, 40, 0.5 / 40)):
    ev = BatchEvaluator(m)
    lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
    q, qd, tau, f = synth_batch(lim, 0, B, N, device="cuda")
    U = B * N
    ms = timed(lambda: ev.rnea(q, qd))
    print("%s  U=%d  rnea      %8.3f ms  %.3e units/s" % (name, U, ms, U / ms * 1e3))
    W = torch.zeros((6, U), dtype=torch.float64, device="cuda")
    ms = timed(lambda: ev.node_eval_ref([m.nframes - 1], 1.0, q, qd, W, f, 0.5))
    print("%s  U=%d  node_eval %8.3f ms  %.3e units/s" % (name, U, ms, U / ms * 1e3))
    ms = timed(lambda: ev.step_rk4(q, qd, tau, f, dt))
    print("%s  U=%d  step      %8.3f ms  %.3e units/s" % (name, U, ms, U / ms * 1e3))
    Uj = U if m.kernel_family.startswith(("chain", "forest")) else 5120
    sl = [t[:, :Uj].contiguous() for t in (q, qd, tau, f)]
    ms = timed(lambda: ev.step_rk4_jvp(*sl, dt), reps=2)
    print("%s  U=%d  step+jac  %8.3f ms  %.3e units/s" % (name, Uj, ms, Uj / ms * 1e3))
