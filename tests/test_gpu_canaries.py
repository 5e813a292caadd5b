"""GPU: out-of-bounds WRITE check without the sanitizer (closed on this pool): every output of the Jacobian entry — states, the
dense Jacobian, the caller workspace — is carved out of one allocation with canary gaps before and after it; the gaps must come
back untouched, for ragged unit counts, all three pipeline kinds (static chain, two-arm forest with coupling, run-time tree) and
a workspace that forces several chunks."""
import ctypes as C

import numpy as np
import pytest

from mpc_fatigue_b200.model import Model, data_urdf

pytestmark = pytest.mark.gpu
GAP = 4096  # doubles on each side of every carved buffer
CANARY = -7.25e300


def _carve(pool, sizes):
    """views of `pool` separated by GAP canary doubles; returns (views, gap slices)"""
    views, gaps, off = [], [], 0
    for sz in sizes:
        gaps.append(slice(off, off + GAP))
        off += GAP
        off = (off + 15) // 16 * 16  # 128-byte alignment of every buffer
        views.append(pool[off:off + sz])
        off += sz
    gaps.append(slice(off, off + GAP))
    return views, gaps, off + GAP


def _models():
    from mpc_fatigue_b200.coupling import box_load_coupling
    m2 = Model.from_urdf(data_urdf("pilz6x2"), armature=1e-2)
    m2.set_coupling(box_load_coupling(m2, mass=30.0))
    return {
        "pilz6": Model.from_urdf(data_urdf("pilz6"), armature=1e-2),
        "pilz6x2_coupled": m2,
        "humanoid37": Model.synthetic("humanoid", 37, seed=7, armature=1e-2),
        "chain9": Model.synthetic("chain", 9, seed=4, armature=1e-2),
    }


@pytest.mark.parametrize("name", ["pilz6", "pilz6x2_coupled", "humanoid37", "chain9"])
@pytest.mark.parametrize("U", [1, 33, 131])
def test_outputs_stay_inside_their_buffers(name, U):
    import torch
    from mpc_fatigue_b200 import _capi
    from mpc_fatigue_b200.synth import synth_batch
    m = _models()[name]
    n = m.n
    lim = {k: m.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
    q, qd, tau, f = synth_batch(lim, 0, U, 1, seed=3, device="cuda")
    # a workspace for ~40 units: U = 131 runs in several chunks (the tree pipeline keeps its unit counters behind the chunk data)
    ws_bytes = int(_capi.lib.mpcf_step_rk4_jvp_workspace_bytes(m.handle, min(U, 40)))
    sizes = [n * U, n * U, n * U, 3 * n * (4 * n + 1) * U, ws_bytes // 8]
    total = _carve(np.empty(0), sizes)[2]
    pool = torch.full((total,), CANARY, dtype=torch.float64, device="cuda")
    (qn, qdn, fn, jac, ws), gaps, _ = _carve(pool, sizes)
    p = lambda t: C.c_void_p(t.data_ptr())
    _capi.check(_capi.lib.mpcf_step_rk4_jvp_ws_batch(m.handle, U, p(q), p(qd), p(tau), p(f), 0.0125, None, p(qn), p(qdn), p(fn), p(jac), p(ws),
                                                     ws_bytes, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    for i, g in enumerate(gaps):
        assert bool((pool[g] == CANARY).all()), (name, U, "canary gap %d was written" % i)
    # and every output element was written (no canary left inside the state and Jacobian buffers)
    for buf, nm in ((qn, "q+"), (qdn, "qd+"), (fn, "f+"), (jac, "jac")):
        assert not bool((buf == CANARY).any()), (name, U, nm, "has unwritten elements")
        assert bool(torch.isfinite(buf).all()), (name, U, nm)
