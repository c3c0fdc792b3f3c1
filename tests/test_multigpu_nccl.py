"""NCCL edge of the batch sharding on real GPUs: one dist.scatter of the mixtures, no collective
inside the iteration, one dist.gather of the separated spectrograms.  Needs >= 2 CUDA devices (skipped otherwise);
the same code runs under gloo on the CPU in tests/test_sharding_gloo.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ssspy_b200 import _lib
    from ssspy_b200.bss import GaussILRMA
    from ssspy_b200.parallel import gather_batch, scatter_batch, separate_sharded, shard_range
    from ssspy_b200.utils.synth import make_batch, make_nmf_init
    N, I, J, K, n_iter = 2, 33, 48, 5, 3
    T, V = make_nmf_init(N, I, J, K, seed=7)
    full = None
    if rank == 0:
        full = torch.from_numpy(make_batch(B, N, I, J, config_id=31).astype(np.complex64)).cuda()
    # plain round trip first: gather(scatter(X)) == X, shards contiguous and balanced
    shard = scatter_batch(full, src=0)
    lo, hi = shard_range(B, rank, world)
    assert shard.is_cuda and shard.shape[0] == hi - lo
    back = gather_batch(shard, B, dst=0)
    n0 = _lib.launch_count()
    Y = separate_sharded(lambda: GaussILRMA(n_basis=K, spatial_algorithm="IP", record_loss=False), full, n_iter,
                         basis=T, activation=V)
    launched = _lib.launch_count() - n0
    ok = True
    if rank == 0:
        ok = bool(torch.equal(back, full))
        ref = GaussILRMA(n_basis=K, spatial_algorithm="IP", record_loss=False)
        ref.chunk_size = B
        Yref = ref(full, n_iter=n_iter, basis=T, activation=V)
        # a mixture's result does not depend on which rank or chunk ran it
        ok = ok and bool(torch.equal(Y, Yref))
    q.put((rank, ok, int(launched), hi - lo))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5, 1])
def test_separate_sharded_over_nccl(B):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(r[1] for r in res)
    # every rank that owns mixtures ran the CUDA path itself
    assert all(r[2] > 0 for r in res if r[3] > 0)
