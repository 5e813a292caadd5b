import sys, os
sys.path.insert(0, os.getcwd())
import torch
from mpc_fatigue_b200.evaluator import BatchEvaluator
from mpc_fatigue_b200.model import Model, data_urdf
from mpc_fatigue_b200.synth import synth_batch
m = Model.from_urdf(data_urdf("pilz6"), armature=1e-2)
ev = BatchEvaluator(m)
lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
B, N = 65536, 100
q, qd, tau, f = synth_batch(lim, 0, B, N, device="cuda")
U = B * N
fr = m.frame_id("prbt_link_5")
W = torch.randn((6, U), dtype=torch.float64, device="cuda")
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
pos = torch.empty((3, U), dtype=torch.float64, device="cuda"); rot = torch.empty((9, U), dtype=torch.float64, device="cuda")
J = torch.empty((36, U), dtype=torch.float64, device="cuda")
o = torch.empty((6, U), dtype=torch.float64, device="cuda")
for name, fn, bytes_ in (("fk", lambda: ev.fk(fr, q, pos, rot), 18 * 8), ("jacobian", lambda: ev.jacobian(fr, q, J), 42 * 8),
                         ("jac_t_wrench", lambda: ev.jac_t_wrench(fr, q, W, o), 18 * 8)):
    ms = timed(fn)
    print("%-14s %.3f ms  %.3e units/s  %.2f TB/s" % (name, ms, U / ms * 1e3, bytes_ * U / ms * 1e-9))

Us = 1 << 20
sl = [t[:, :Us].contiguous() for t in (q, qd, tau)]
ms = timed(lambda: ev.fd_derivs(*sl))
print("%-14s %.3f ms  %.3e units/s  (U = %d)" % ("fd_derivs", ms, Us / ms * 1e3, Us))
q0, qd0, f0 = (t[:, :B].contiguous() for t in (q, 0.2 * qd, f))
ms = timed(lambda: ev.rollout_rk4(q0, qd0, f0, 0.3 * tau, N, 0.02), reps=3)
print("%-14s %.3f ms  %.3e rollout-steps/s  (B = %d scenarios x N = %d steps, one thread per scenario)" % ("rollout", ms, U / ms * 1e3, B, N))
ms = timed(lambda: ev.aba(q, qd, tau))
print("%-14s %.3f ms  %.3e units/s" % ("fwd dynamics", ms, U / ms * 1e3))
ms = timed(lambda: ev.step_rk4(q, qd, tau, f, 0.02))
print("%-14s %.3f ms  %.3e units/s" % ("step_rk4", ms, U / ms * 1e3))
