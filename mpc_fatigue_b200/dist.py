"""Scenario sharding across the GPUs of one box.

Units are independent, so evaluation needs no communication: rank r owns a contiguous block of
scenarios (all N nodes of a scenario stay on one GPU).  The only collective is one all-gather of the
per-scenario (cost, residual) rows, which the reduction kernel writes directly into the send buffer.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(B_total: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the scenarios owned by `rank`; block partition, remainder to the first ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(B_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allgather_rows(local: torch.Tensor, counts: list[int] | None = None, group=None) -> torch.Tensor:
    """All-gather `[R, B_local]` rows into `[R, B_total]` (scenario order = rank order).

    NCCL over NVLink on GPUs; gloo in the CPU tests.  Equal shards use one all_gather_into_tensor; ragged
    shards are padded to the largest shard and trimmed.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    R, Bl = local.shape
    if counts is None:
        counts = [Bl] * world
    Bmax = max(counts)
    send = local if Bl == Bmax else torch.nn.functional.pad(local, (0, Bmax - Bl))
    send = send.contiguous()
    recv = torch.empty((world, R, Bmax), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)  # flat views: accepted by NCCL and gloo alike
    if all(c == Bmax for c in counts):
        return recv.permute(1, 0, 2).reshape(R, world * Bmax)
    return torch.cat([recv[r, :, : counts[r]] for r in range(world)], dim=1)


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> int | None:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that the pinned staging buffers it
    allocates afterwards (first touch) and the copy-engine traffic into them stay on that node.  One process per GPU:
    without this the ranks of a multi-GPU run share whatever node the launcher left them on.
    Returns the node, or None when the topology cannot be read (nothing is changed then)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = _parse_cpulist(fh.read()) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, AttributeError, ValueError, RuntimeError, AssertionError):  # no sysfs entry / no driver / old torch
        return None
