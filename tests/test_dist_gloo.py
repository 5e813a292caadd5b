"""world_size-2 gloo test of the multi-GPU host path (runs on CPU): block sharding + the all-gather of the
per-scenario (cost, residual) rows reproduces the unsharded order; ragged shards too."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_fatigue_b200.dist import allgather_rows, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard_range(B, rank, world)
        # per-scenario rows are a pure function of the global scenario index (stand-in for the reduction kernel)
        idx = torch.arange(a, b, dtype=torch.float64)
        local = torch.stack([idx * 2.0 + 1.0, idx ** 2, -idx, idx + 0.5])
        counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
        full = allgather_rows(local, counts)
        ref_idx = torch.arange(B, dtype=torch.float64)
        ref = torch.stack([ref_idx * 2.0 + 1.0, ref_idx ** 2, -ref_idx, ref_idx + 0.5])
        ok = tuple(full.shape) == (4, B) and torch.equal(full, ref)
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(float(t.item()))
    finally:
        dist.destroy_process_group()


def _run(B):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1.0


def test_allgather_equal_shards():
    _run(64)


def test_allgather_ragged_shards():
    _run(37)


def test_single_process_is_identity():
    x = torch.arange(12, dtype=torch.float64).reshape(4, 3)
    assert allgather_rows(x) is x
