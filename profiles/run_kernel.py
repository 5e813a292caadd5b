#!/usr/bin/env python3
"""Profiling driver: launches one hot-path kernel a few times on the C2 workload shape (smaller B by default).
    python profiles/run_kernel.py jvp|step|rnea|node [B] [reps] [pilz6|pilz6x2|pilz6x2c|humanoid37] [N]
Used under ncu (see profiles/README.md); never a source of bench numbers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mpc_fatigue_b200.evaluator import BatchEvaluator
from mpc_fatigue_b200.model import Model, data_urdf
from mpc_fatigue_b200.synth import synth_batch

which = sys.argv[1] if len(sys.argv) > 1 else "jvp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mname = sys.argv[4] if len(sys.argv) > 4 else "pilz6"
N = int(sys.argv[5]) if len(sys.argv) > 5 else 100
dt = 0.02
dev = torch.device("cuda", 0)
if mname == "humanoid37":
    m = Model.synthetic("humanoid", 37, seed=7, armature=1e-2)
elif mname == "pilz6x2c":  # config C3: both arms with the coupled fatigue (box load split)
    from mpc_fatigue_b200.coupling import box_load_coupling
    m = Model.from_urdf(data_urdf("pilz6x2"), armature=1e-2)
    m.set_coupling(box_load_coupling(m))
else:
    m = Model.from_urdf(data_urdf(mname), armature=1e-2)
ev = BatchEvaluator(m, dev)
lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
q, qd, tau, f = synth_batch(lim, 0, B, N, device=dev)
U = B * N
out = tuple(torch.empty_like(q) for _ in range(3))
jac = torch.empty((3 * m.n, 4 * m.n + 1, U), dtype=torch.float64, device=dev) if which == "jvp" else None
W = torch.zeros((6, U), dtype=torch.float64, device=dev)
fr = m.frame_id("prbt_link_5") if mname == "pilz6" else m.nframes - 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps + 1):
    if r == 1:
        e0.record()
    if which == "jvp":
        ev.step_rk4_jvp(q, qd, tau, f, dt, out=out, jac=jac)
    elif which == "step":
        ev.step_rk4(q, qd, tau, f, dt, out=out)
    elif which == "rnea":
        ev.rnea(q, qd, None, out=out[0])
    elif which == "node":
        ev.node_eval_ref([fr], -1.0, q, qd, W, f, dt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("%s: U=%d  %.3f ms/launch  %.3e units/s" % (which, U, ms, U / ms * 1e3))
if which == "nodejvp":
    import time
    for r in range(reps + 1):
        if r == 1:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        ev.node_eval_ref_jvp([fr], -1.0, q, qd, W)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print("nodejvp: U=%d  %.3f ms/launch  %.3e units/s" % (U, ms, U / ms * 1e3))
    ev.rnea_derivs(q, qd); torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(reps):
        ev.rnea_derivs(q, qd)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print("rnea_derivs: U=%d  %.3f ms/launch  %.3e units/s" % (U, ms, U / ms * 1e3))
