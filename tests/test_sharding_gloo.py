"""world_size-2 gloo test (CPU) of the batch scatter / gather used for multi-GPU runs: shards are
contiguous, ragged batches work, and gather(scatter(X)) == X; no collective is needed in between."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssspy_b200.parallel import gather_batch, scatter_batch, shard_range
    rng = np.random.default_rng(0)
    full = None
    if rank == 0:
        full = torch.from_numpy((rng.standard_normal((B, 2, 5, 7)) + 1j * rng.standard_normal((B, 2, 5, 7))).astype(np.complex64))
    shard = scatter_batch(full, src=0)
    lo, hi = shard_range(B, rank, world)
    assert shard.shape[0] == hi - lo
    # per-mixture independent "work": scale each mixture by its global index + 1
    idx = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1, 1) + 1
    out = gather_batch(shard * idx, B, dst=0)
    if rank == 0:
        want = full * (torch.arange(B, dtype=torch.float32).view(-1, 1, 1, 1) + 1)
        q.put(bool(torch.allclose(out, want)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5, 1])
def test_scatter_gather_world2(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
