"""Multi-GPU batch sharding.

Mixtures are independent, so the only sensible parallelism is data parallel over the batch: rank g
owns mixtures ``[g*B/G, (g+1)*B/G)``; one scatter of the inputs before the loop and one gather of
the outputs after it (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests); NO
collective inside the iteration.  A single mixture is never split across GPUs (that would put an
all-reduce into every iteration).
"""
import torch
import torch.distributed as dist


def shard_range(n_batch, rank, world_size):
    """Contiguous, balanced partition of range(n_batch): the first ``n_batch % world_size`` ranks get
    one extra mixture."""
    base, extra = divmod(n_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scatter_batch(full, src=0, group=None):
    """``full`` (B, ...) on ``src`` (None elsewhere) -> this rank's shard.  One collective."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = [None]
    if rank == src:
        meta = [(tuple(full.shape), full.dtype, full.device.type)]
    dist.broadcast_object_list(meta, src=src, group=group)
    shape, dtype, devtype = meta[0]
    device = torch.device("cuda", torch.cuda.current_device()) if devtype == "cuda" else torch.device("cpu")
    lo, hi = shard_range(shape[0], rank, world)
    out = torch.empty((hi - lo,) + tuple(shape[1:]), dtype=dtype, device=device)
    is_c = dtype.is_complex
    view = torch.view_as_real(out) if is_c else out
    if rank == src:
        chunks = []
        for r in range(world):
            a, b = shard_range(shape[0], r, world)
            c = full[a:b].contiguous()
            chunks.append(torch.view_as_real(c) if is_c else c)
        # ragged shards: point-to-point sends keep it to one message per rank
        reqs = [dist.isend(chunks[r], dst=r, group=group) for r in range(world)
                if r != src and chunks[r].shape[0] > 0]
        view.copy_(chunks[src])
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(view, src=src, group=group)
    return out


def gather_batch(shard, n_batch, dst=0, group=None):
    """Inverse of :func:`scatter_batch`: returns the (n_batch, ...) tensor on ``dst``, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    is_c = shard.dtype.is_complex
    view = torch.view_as_real(shard.contiguous()) if is_c else shard.contiguous()
    if rank != dst:
        if shard.shape[0] > 0:
            dist.send(view, dst=dst, group=group)
        return None
    full = torch.empty((n_batch,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
    fview = torch.view_as_real(full) if is_c else full
    for r in range(world):
        a, b = shard_range(n_batch, r, world)
        if r == dst:
            fview[a:b].copy_(view)
        elif b > a:
            dist.recv(fview[a:b], src=r, group=group)
    return full


def separate_sharded(make_separator, X_full, n_iter, src=0, group=None, **state):
    """Scatter ``X_full`` (B, N, I, J) from ``src``, run ``make_separator()(shard, n_iter)`` on every
    rank with no communication, gather the outputs on ``src``."""
    n_batch = [None]
    if dist.get_rank(group) == src:
        n_batch = [int(X_full.shape[0])]
    dist.broadcast_object_list(n_batch, src=src, group=group)
    shard = scatter_batch(X_full, src=src, group=group)
    if shard.shape[0] == 0:
        Y = shard
    else:
        Y = make_separator()(shard, n_iter=n_iter, **state)
    return gather_batch(Y, n_batch[0], dst=src, group=group)
