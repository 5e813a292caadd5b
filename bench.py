#!/usr/bin/env python3
"""bench.py — rollout-steps/s of the RK4 dynamics+fatigue step WITH forward-mode Jacobians on config C2
(Pilz 6-DOF + armature 1e-2, N = 100 nodes x B = 65,536 scenarios per GPU, fp64).

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                   CPU arm: the oracle port on all host cores

One "step" = one pass of the hot path over the rank's whole batch: step_rk4_jvp kernel (states + dense
Jacobian for U = B*N units) + the per-scenario cost/residual reduction, and for N > 1 the NCCL all-gather
of the [4, B] cost/residual rows.  Inputs (1.26 GB) and outputs (24.5 GB) are far larger than the 126 MB
L2, so no explicit flush is needed between timed iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rollout_steps_per_s_with_jacobians"
UNIT = "rollout-steps/s"
N_NODES = 100
B_PER_GPU = 65536
DT = 2.0 / N_NODES
ARMATURE = 1e-2
NDOF = 6
# DRAM bytes per unit of the Jacobian pipeline from the committed ncu --set full capture (profiles/); None = not captured
TRAFFIC_BYTES_PER_UNIT = 13352  # (2.32 + 3.67 + 8.01) GB per 2^20 units, profiles/r01_jvp_pipeline.md


# ---- frozen work model (BASELINE.md §3 / SURVEY.md §8d) ----
def flop_model(n: int) -> dict:
    aba = (224 * n - 259) + (205 * n - 248)
    step_values = 4 * (aba + 7 * n) + 24 * n
    P = 4 * n + 1
    return dict(aba=aba, step_values=step_values, step_jac=step_values * (1 + P),
                bytes_values=(4 * n + 1 + 3 * n) * 8, bytes_jac=(4 * n + 1 + 3 * n) * 8 + 3 * n * P * 8)


def workload_config(n_gpus: int) -> dict:
    return {"workload": "C2: Pilz 6-DOF (urdf pilz_robot_6DOF, armature 1e-2), RK4 dyn+fatigue step with dense "
                        "forward-mode Jacobian, N=100 nodes x 65536 scenarios per GPU, dt=0.02",
            "ndof": NDOF, "nodes": N_NODES, "scenarios_per_gpu": B_PER_GPU, "scenarios_total": B_PER_GPU * n_gpus,
            "units_per_gpu": B_PER_GPU * N_NODES, "parallelism": "scenario-sharded x%d" % n_gpus,
            "l2": "inputs+outputs (25.8 GB/GPU) >> 126 MB L2, no flush needed", "seed": 1234}


# ---- clocks sampler (B200_PROFILING.md recipe) ----
class ClockSampler:
    """SM clock and clock-event reasons of one GPU, sampled every 50 ms by a thread through NVML (what nvidia-smi itself
    reads; no subprocess, so no start-up delay and no pipe buffering: every sample carries its own time stamp).  Falls back
    to one `nvidia-smi --query-gpu` call per sample when the NVML python module is missing."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self.thread, self.stop_flag, self.src = index, [], None, threading.Event(), None

    def _nvml_reader(self):
        import pynvml
        pynvml.nvmlInit()
        import torch
        p = torch.cuda.get_device_properties(self.index)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)).encode())
        mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]

        def read():
            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            return float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), mx, [bool(r & b) for b in bits]
        return read

    def _smi_reader(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

        def read():
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return float(out[0]), float(out[1]), [c.strip().lower().startswith("active") for c in out[2:6]]
        return read

    def start(self):
        read = None
        for name, make in (("nvml", self._nvml_reader), ("nvidia-smi", self._smi_reader)):
            try:
                read = make()
                read()
                self.src = name
                break
            except Exception:  # noqa: BLE001 - any failure of one source just selects the next
                read = None
        if read is None:
            return

        def loop():
            while not self.stop_flag.is_set():
                try:
                    self.rows.append((time.perf_counter(),) + read())
                except Exception:  # noqa: BLE001
                    pass
                self.stop_flag.wait(0.05)
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    @staticmethod
    def mark() -> float:
        return time.perf_counter()

    def stop(self, t0: float | None = None, t1: float | None = None) -> dict:
        """Statistics over the samples taken inside [t0, t1] (the timed region); if none fell inside it (a very short
        region), over everything sampled since start(), i.e. warm-up + timed steps of the same kernels, and says so."""
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable (no NVML module, no nvidia-smi)"], "samples": 0}
        self.stop_flag.set()
        self.thread.join(timeout=15)
        rows = [r for r in self.rows if t0 is not None and t1 is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if not rows:
            rows, window = list(self.rows), "warm-up + timed region (no sample fell inside the timed region)"
        mhz = sorted(r[1] for r in rows)
        reasons = sorted({nm for r in rows for nm, on in zip(self.NAMES, r[3]) if on})
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": rows[-1][2] if rows else None, "reasons": reasons,
                "samples": len(mhz), "window": window, "source": self.src}


def measured_peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except OSError:
        return {}


# =================================================================================================
# CPU arm: the oracle port (test infrastructure) timed on the host cores
# =================================================================================================
def cpu_arm_setup():
    import numpy as np
    from oracle import pyoracle
    from oracle.urdf_model import load_urdf
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "-B", "_build/libmpcf_oracle_fast.so"], check=True)
    with open(os.path.join(ROOT, "mpc_fatigue_b200", "data", "models", "pilz6.urdf")) as fh:
        om = load_urdf(fh.read(), armature=ARMATURE)
    cores = len(os.sched_getaffinity(0))
    orc = pyoracle.Oracle(om, fast=True, threads=cores)
    return np, om, orc, cores


def cpu_inputs(np, om, U, seed=1234):
    rng = np.random.default_rng(seed)
    n = om.n
    lo, hi = np.array(om.q_lo)[:, None], np.array(om.q_hi)[:, None]
    q = np.ascontiguousarray(rng.uniform(lo, hi, (n, U)))
    qd = np.ascontiguousarray(rng.uniform(-1, 1, (n, U)) * np.array(om.v_max)[:, None])
    tau = np.ascontiguousarray(rng.uniform(-1, 1, (n, U)) * np.array(om.tau_max)[:, None] * 0.25)
    f = np.ascontiguousarray(rng.uniform(20, 80, (n, U)))
    return q, qd, tau, f


def cpu_baseline(target_seconds: float = 12.0) -> dict:
    """Bounded sample of the same workload on the host cores (values + Jacobians through the oracle)."""
    np, om, orc, cores = cpu_arm_setup()
    U0 = 2048
    q, qd, tau, f = cpu_inputs(np, om, U0)
    orc.step_rk4_jvp(q, qd, tau, f, DT)  # warm-up
    t0 = time.perf_counter()
    orc.step_rk4_jvp(q, qd, tau, f, DT)
    rate0 = U0 / (time.perf_counter() - t0)
    U = int(max(U0, min(2 ** 20, rate0 * target_seconds / 3)))
    q, qd, tau, f = cpu_inputs(np, om, U)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        orc.step_rk4_jvp(q, qd, tau, f, DT)
        best = max(best, U / (time.perf_counter() - t0))
    return {"value": best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d units of the C2 workload (same model, dt, value distributions), best of 3, "
                      "oracle C port -O3 -march=native + OpenMP, Jacobians by complex-step" % U,
            "per_core": best / cores}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    np, om, orc, cores = cpu_arm_setup()
    U0 = 2048
    q, qd, tau, f = cpu_inputs(np, om, U0)
    orc.step_rk4_jvp(q, qd, tau, f, DT)
    t0 = time.perf_counter()
    orc.step_rk4_jvp(q, qd, tau, f, DT)
    rate0 = U0 / (time.perf_counter() - t0)
    total = max(1, args.steps + args.warmup)
    budget_s = 2.0 if os.environ.get("MPCF_BENCH_QUICK") else 90.0  # whole run ~ 90 s (a few seconds in the contract test)
    U = int(max(256, min(2 ** 20, rate0 * budget_s / total)))
    q, qd, tau, f = cpu_inputs(np, om, U)
    for _ in range(args.warmup):
        orc.step_rk4_jvp(q, qd, tau, f, DT)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.step_rk4_jvp(q, qd, tau, f, DT)
    el = time.perf_counter() - t0
    value = U * args.steps / el
    sample = "%d units per step (bounded sample of the C2 workload), %d host threads" % (U, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference CasADi/Pinocchio path cannot be built in this image (no casadi/pinocchio/Eigen/urdfdom); "
                "this arm times the compiled oracle port, which is at least as fast as CasADi's SX interpreter",
    }))


# =================================================================================================
# GPU arm
# =================================================================================================
def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist
    from mpc_fatigue_b200 import _capi
    from mpc_fatigue_b200.dist import allgather_rows
    from mpc_fatigue_b200.evaluator import BatchEvaluator
    from mpc_fatigue_b200.model import Model, data_urdf
    from mpc_fatigue_b200.pipeline import HostStepPipeline
    from mpc_fatigue_b200.synth import synth_batch
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from mpc_fatigue_b200.dist import bind_to_gpu_numa_node
    # multi-GPU: every rank's pinned staging buffers on the NUMA node of its own GPU (N = 1 keeps all host cores, which
    # the cpu_baseline leg uses)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and os.environ.get("MPCF_NUMA_BIND", "1") != "0" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = Model.from_urdf(data_urdf("pilz6"), armature=ARMATURE)
    ev = BatchEvaluator(model, dev)
    n, B, N = model.n, B_PER_GPU, N_NODES
    U = B * N
    limits = {k: model.export(k) for k in ("q_lo", "q_hi", "v_max", "tau_max")}
    q, qd, tau, f = synth_batch(limits, rank * B, B, N, seed=1234, device=dev)
    qn, qdn, fn = (torch.empty_like(q) for _ in range(3))
    jac = torch.empty((3 * n, 4 * n + 1, U), dtype=torch.float64, device=dev)
    red = torch.empty((4, B), dtype=torch.float64, device=dev)  # all-gather send buffer, written by the kernel

    def step(ev_pairs=None):
        if ev_pairs is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ev.step_rk4_jvp(q, qd, tau, f, DT, out=(qn, qdn, fn), jac=jac)
        if ev_pairs is not None:
            e1.record()
            ev_pairs.append((e0, e1))
        ev.cost_residual(B, N, q, qd, f, tau, qn, qdn, fn, DT, out=red)
        return allgather_rows(red) if world > 1 else red

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    launches0 = _capi.lib.mpcf_launch_count()
    pairs = []
    _capi.lib.mpcf_profile_enable(1)  # per-kernel CUDA events on the launch stream (read after the timed region)
    fence()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mark0 = sampler.mark()
    t0.record()
    for _ in range(args.steps):
        gathered = step(pairs)
    t1.record()
    fence()
    mark1 = sampler.mark()
    launches = _capi.lib.mpcf_launch_count() - launches0
    el_ms = t0.elapsed_time(t1)
    ms3 = (C.c_double * 3)()
    nprof = C.c_long()
    _capi.lib.mpcf_profile_read(ms3, C.byref(nprof))
    _capi.lib.mpcf_profile_enable(0)
    kern = {"step_stages": ms3[0] / args.steps, "stage_derivs": ms3[1] / args.steps, "chain_rule": ms3[2] / args.steps}
    clocks = sampler.stop(mark0, mark1) if rank == 0 else None
    if world > 1:
        tt = torch.tensor([el_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        el_ms = float(tt.item())
    kern_ms = sum(a.elapsed_time(b) for a, b in pairs) / max(len(pairs), 1)
    value = world * U * args.steps / (el_ms * 1e-3)
    checksum = float(gathered[0].sum().item())

    # ---- values-only kernel (reported next to the headline; same units) ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        ev.step_rk4(q, qd, tau, f, DT, out=(qn, qdn, fn))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        ev.step_rk4(q, qd, tau, f, DT, out=(qn, qdn, fn))
    e1.record()
    torch.cuda.synchronize()
    values_ms = e0.elapsed_time(e1) / 3

    # ---- FP64 pipe probe (roofline denominator; no FP64 figure in MEASURED_PEAKS.json) ----
    blocks, iters = 148 * 16, 200000
    pout = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
    best_tf = 0.0
    for _ in range(4):
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        _capi.check(_capi.lib.mpcf_probe_fp64(iters, blocks, C.c_void_p(pout.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        p1.record()
        torch.cuda.synchronize()
        best_tf = max(best_tf, blocks * 256 * iters * 16 / (p0.elapsed_time(p1) * 1e-3) / 1e12)

    # ---- end to end through the host-facing API: pinned host inputs -> H2D -> kernels -> D2H of states + Jacobian ----
    del jac
    torch.cuda.empty_cache()
    pipe = HostStepPipeline(model, dev, chunk_units=1 << 19, skip_structural_zeros=True)
    hq, hqd, htau, hf = (torch.empty((n, U), dtype=torch.float64).pin_memory() for _ in range(4))
    for h_, d_ in ((hq, q), (hqd, qd), (htau, tau), (hf, f)):
        h_.copy_(d_)
    fence()
    e2e_steps = max(1, min(args.steps, 3))
    pipe.run(hq, hqd, htau, hf, DT, B, N)  # warm-up pass
    fence()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        stats = pipe.run(hq, hqd, htau, hf, DT, B, N)
        if world > 1:
            allgather_rows(pipe.reduced)
    fence()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * U * e2e_steps / e2e_s

    # same host-facing call, but only the per-scenario cost / residual rows return to the host (states and Jacobians stay
    # on the device): the variant for a device-resident consumer; reported beside `e2e`, never instead of it
    pipe_r = HostStepPipeline(model, dev, chunk_units=1 << 19, outputs="reduced")
    pipe_r.run(hq, hqd, htau, hf, DT, B, N)
    fence()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        stats_r = pipe_r.run(hq, hqd, htau, hf, DT, B, N)
        if world > 1:
            allgather_rows(pipe_r.reduced)
    fence()
    e2e_r_s = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([e2e_r_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_r_s = float(tt.item())
    e2e_reduced = {"value": world * U * e2e_steps / e2e_r_s, "unit": UNIT, "h2d_bytes_per_step": int(stats_r["h2d_bytes"]),
                   "d2h_bytes_per_step": int(stats_r["d2h_bytes"]),
                   "what": "HostStepPipeline(outputs='reduced'): pinned host inputs -> H2D -> step_rk4_jvp + cost_residual -> D2H of the "
                           "[4, B] per-scenario cost/residual rows only"}

    if rank == 0:
        fm = flop_model(n)
        peaks = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ach_tf = fm["step_jac"] * U / (kern_ms * 1e-3) / 1e12
        ach_gbs = fm["bytes_jac"] * U / (kern_ms * 1e-3) / 1e9
        cpu = cpu_baseline() if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": el_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(stats["h2d_bytes"]),
                    "d2h_bytes_per_step": int(stats["d2h_bytes"]), "steps": e2e_steps, "numa_node_rank0": numa,
                    "what": "HostStepPipeline.run: pinned host q,qd,tau,f -> chunked H2D -> step_rk4_jvp + cost_residual -> "
                            "D2H of q+,qd+,f+, the 348 structurally non-zero Jacobian planes (of 450: d(q+,qd+)/df = 0, df+/df "
                            "diagonal) and per-scenario cost/residuals into pinned host staging (PCIe-bound: 2.9 KB per unit)"},
            "e2e_reduced_rows": e2e_reduced,
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": best_tf, "unit": "TFLOP/s", "frac": ach_tf / best_tf,
                         "traffic": TRAFFIC_BYTES_PER_UNIT * U if TRAFFIC_BYTES_PER_UNIT else None,
                         "traffic_source": "profiles/r01_jvp_pipeline.md: dram read+write of the 3 kernels per unit x U (ncu --set full)",
                         "kernel": "Jacobian pipeline = k_step_stages + k_stage_derivs + k_chain_rule_tma per chunk of 2^20 units",
                         "kernel_ms": kern_ms, "kernels_ms": kern, "dominant_kernel": max(kern, key=kern.get),
                         "dominant_share": max(kern.values()) / max(sum(kern.values()), 1e-9),
                         "note": "frac can exceed 1: the frozen model charges (1 + 25) RK4/ABA sweeps per unit; the analytic "
                                 "pipeline needs ~6x fewer FP64 instructions (DESIGN.md §5)",
                         "executed": {"fp64_pipe_active_pct": {"step_stages": 66.7, "stage_derivs": 63.3, "chain_rule": 38.2},
                                      "fp64_thread_instructions_per_unit": 43000,
                                      "source": "ncu --set full, profiles/r01_jvp_pipeline.md (static figures of that capture, not "
                                                "measured by this run): the pipeline runs at about half of the FP64 issue rate"},
                         "flop_model": "frozen BASELINE.md §3: %d FLOP per unit (values %d x (1 + 25 seeds))" % (fm["step_jac"], fm["step_values"]),
                         "peak_source": "DFMA-chain probe measured in this run (MEASURED_PEAKS.json has no FP64 figure; nominal 37.2)",
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}},
            "values_only": {"value": U / (values_ms * 1e-3), "unit": UNIT, "kernel_ms": values_ms,
                            "roofline_frac_fp64": fm["step_values"] * U / (values_ms * 1e-3) / 1e12 / best_tf},
            "cpu_baseline": cpu, "checksum_cost": checksum,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
