"""HostStepPipeline — the host-facing batched call: HOST (pinned) inputs in, HOST outputs out.

This is the end-to-end path a CPU-side NLP solver would use in place of the reference's per-node
`Function(q=..., ...)` calls (python/Pilz_6_DOF/force_optimization_pilz_6DOF.py:259-274): the scenario axis
is cut into chunks; for each chunk the input planes are copied host->device, the RK4+Jacobian kernel and the
per-scenario cost/residual reduction run, and states + dense Jacobian + reduced rows are copied
device->host into pinned staging.  Three streams (H2D, compute, D2H) and two buffer slots overlap the
copies of neighbouring chunks with compute.  What goes back is first packed on the device (mpcf_gather_planes: states + the
selected Jacobian planes -> one contiguous buffer), so every chunk returns in ONE device->host copy.  Host layout = device layout: `[component, N, B]` planes with
node-major units, so a scenario chunk of every plane is one pitched DMA (mpcf_memcpy2d_async).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from .evaluator import BatchEvaluator
from .model import Model


class HostStepPipeline:
    def __init__(self, model: Model, device, chunk_units: int = 1 << 19, with_jacobian: bool = True, skip_structural_zeros: bool = False,
                 outputs: str = "all"):
        """outputs = "all": states + Jacobian planes + per-scenario rows come back to the host (what a host-side NLP solver needs).
        outputs = "reduced": only the per-scenario (cost, defect, torque-bound, fatigue-bound) rows come back; states and
        Jacobians are computed but stay on the device (what the multi-GPU path gathers)."""
        if outputs not in ("all", "reduced"):
            raise ValueError("outputs must be 'all' or 'reduced'")
        self.outputs = outputs
        self.model, self.dev = model, torch.device(device)
        self.ev = BatchEvaluator(model, self.dev)
        self.chunk_units = int(chunk_units)
        self.with_jacobian = with_jacobian
        self.n = model.n
        self._shape = None
        self.reduced = None
        # Device output planes per unit: q+ (n), qd+ (n), f+ (n), then the Jacobian planes r*(4n+1)+c.  With
        # skip_structural_zeros only the planes that can be non-zero travel to the host: d(q+, qd+)/df = 0 and df+/df is
        # diagonal, i.e. 348 of 450 Jacobian planes for n = 6.  `plane_map[k]` = device plane of host row k.
        n = self.n
        P = 4 * n + 1
        planes = list(range(3 * n))
        if with_jacobian:
            for r in range(3 * n):
                for c in range(P):
                    fcol = 3 * n <= c < 4 * n
                    if skip_structural_zeros and fcol and not (r >= 2 * n and c - 3 * n == r - 2 * n):
                        continue
                    planes.append(3 * n + r * P + c)
        self.plane_map = planes
        self.d_map = torch.tensor(planes, dtype=torch.int32, device=self.dev)
        self.identity_map = planes == list(range(len(planes)))  # nothing skipped: the output buffer already is the packed form

    def _alloc(self, B: int, N: int):
        n = self.n
        Bc = max(1, min(B, self.chunk_units // N))
        if self._shape == (B, N, Bc):
            return
        Uc = Bc * N
        out_rows = 3 * n + (3 * n * (4 * n + 1) if self.with_jacobian else 0)
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.d_in = [torch.empty((4 * n, Uc), **f64) for _ in range(2)]
        self.d_out = [torch.empty((out_rows, Uc), **f64) for _ in range(2)]
        self.d_red = [torch.empty((4, Bc), **f64) for _ in range(2)]
        self.d_pack = None if (self.identity_map or self.outputs != "all") else [torch.empty((len(self.plane_map), Uc), **f64) for _ in range(2)]
        self.h_out = [torch.empty((len(self.plane_map), Uc), dtype=torch.float64).pin_memory() for _ in range(2)]
        self.h_red = torch.empty((4, B), dtype=torch.float64).pin_memory()
        self.reduced = torch.empty((4, B), **f64)
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self._shape = (B, N, Bc)

    def run(self, hq, hqd, htau, hf, dt: float, B: int, N: int, consume=None) -> dict:
        """hq.. : pinned host float64 tensors [n, N*B] (node-major units).  `consume(view, b0, b1)` is called for
        every finished chunk with the pinned staging view [rows, N, b1-b0] (host row k = device plane `plane_map[k]`:
        q+, qd+, f+, then the Jacobian planes r*(4n+1)+c); without it the staging buffers are simply overwritten."""
        n = self.n
        for t in (hq, hqd, htau, hf):
            if not (t.dtype == torch.float64 and t.is_pinned() and t.is_contiguous() and tuple(t.shape) == (n, N * B)):
                raise ValueError("host inputs must be pinned contiguous float64 [n, N*B] tensors")
        self._alloc(B, N)
        _, _, Bc = self._shape
        h2d = d2h = 0
        ev_cmp = [None, None]
        ev_out = [None, None]
        pending = []
        rows = self.d_out[0].shape[0]
        hrows = len(self.plane_map)
        cur = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event()
        start.record(cur)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_event(start)

        def hand_over(item):
            sl, a0, a1, e = item
            e.synchronize()
            consume(self.h_out[sl].reshape(-1)[:hrows * (a1 - a0) * N].view(hrows, N, a1 - a0), a0, a1)

        with torch.cuda.device(self.dev):
            for ci, b0 in enumerate(range(0, B, Bc)):
                b1 = min(B, b0 + Bc)
                bc = b1 - b0
                uc = bc * N
                slot = ci % 2
                d_in, d_out, d_red = self.d_in[slot], self.d_out[slot], self.d_red[slot]
                # ---- H2D: scenario slice [b0, b1) of every input plane, one pitched DMA per array ----
                if ev_cmp[slot] is not None:
                    self.s_in.wait_event(ev_cmp[slot])  # the slot's previous compute has consumed d_in
                for k, h in enumerate((hq, hqd, htau, hf)):
                    dst = d_in.data_ptr() + k * n * uc * 8
                    _capi.check(_capi.lib.mpcf_memcpy2d_async(C.c_void_p(dst), bc * 8, C.c_void_p(h.data_ptr() + b0 * 8), B * 8,
                                                              bc * 8, n * N, 1, C.c_void_p(self.s_in.cuda_stream)))
                    h2d += n * uc * 8
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
                # ---- compute ----
                self.s_cmp.wait_event(ev_in)
                if ev_out[slot] is not None:
                    self.s_cmp.wait_event(ev_out[slot])  # the slot's previous D2H has drained d_out
                with torch.cuda.stream(self.s_cmp):
                    fl = d_in.reshape(-1)
                    q, qd, tau, f = (fl[k * n * uc:(k + 1) * n * uc].view(n, uc) for k in range(4))
                    fo = d_out.reshape(-1)
                    qn, qdn, fn = (fo[k * n * uc:(k + 1) * n * uc].view(n, uc) for k in range(3))
                    if self.with_jacobian:
                        jac = fo[3 * n * uc:rows * uc].view(3 * n, 4 * n + 1, uc)
                        self.ev.step_rk4_jvp(q, qd, tau, f, dt, out=(qn, qdn, fn), jac=jac)
                    else:
                        self.ev.step_rk4(q, qd, tau, f, dt, out=(qn, qdn, fn))
                    red = d_red.reshape(-1)[:4 * bc].view(4, bc)
                    self.ev.cost_residual(bc, N, q, qd, f, tau, qn, qdn, fn, dt, out=red)
                    self.reduced[:, b0:b1].copy_(red)
                ev_cmp[slot] = torch.cuda.Event()
                ev_cmp[slot].record(self.s_cmp)
                # ---- D2H of the chunk's outputs (contiguous) ----
                self.s_out.wait_event(ev_cmp[slot])
                if self.outputs == "all":
                    if consume is not None and len(pending) == 2:  # the slot about to be overwritten goes to the consumer first
                        hand_over(pending.pop(0))
                    with torch.cuda.stream(self.s_out):
                        hflat = self.h_out[slot].reshape(-1)
                        if self.d_pack is None:
                            packed = d_out.reshape(-1)
                        else:  # pack the planes that travel into one contiguous buffer (device copy kernel), then ONE D2H
                            packed = self.d_pack[slot].reshape(-1)
                            _capi.check(_capi.lib.mpcf_gather_planes(C.c_void_p(d_out.data_ptr()), uc, C.c_void_p(self.d_map.data_ptr()), hrows, uc,
                                                                     C.c_void_p(packed.data_ptr()), C.c_void_p(self.s_out.cuda_stream)))
                        hflat[:hrows * uc].copy_(packed[:hrows * uc], non_blocking=True)
                    d2h += hrows * uc * 8
                ev_out[slot] = torch.cuda.Event()
                ev_out[slot].record(self.s_out)
                if consume is not None and self.outputs == "all":
                    pending.append((slot, b0, b1, ev_out[slot]))
            for item in pending:
                hand_over(item)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_stream(self.s_cmp)
                self.h_red.copy_(self.reduced, non_blocking=True)
            d2h += 4 * B * 8
        for s in (self.s_in, self.s_cmp, self.s_out):
            cur.wait_stream(s)
        cur.synchronize()
        return {"h2d_bytes": h2d, "d2h_bytes": d2h, "chunks": (B + Bc - 1) // Bc}
