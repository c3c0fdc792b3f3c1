"""Multi-GPU batch sharding.

Mixtures are independent, so the only sensible parallelism is data parallel over the batch: rank g
owns mixtures ``[g*B/G, (g+1)*B/G)``; one ``dist.scatter`` of the inputs before the loop and one ``dist.gather``
of the outputs after it (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests); NO
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


def _per_rank(n_batch, world_size):
    return -(-n_batch // world_size)


def _window(n_batch, rank, world_size):
    """Shards travel as equally sized windows of ``per = ceil(B / G)`` mixtures so that ONE scatter / gather
    collective moves them: rank r's window starts at ``min(lo, B - per)`` and its shard is rows
    ``[lo - start, hi - start)`` of it (a short shard carries rows of its neighbour, which are ignored)."""
    per = _per_rank(n_batch, world_size)
    lo, hi = shard_range(n_batch, rank, world_size)
    start = max(min(lo, n_batch - per), 0)
    return start, per, lo - start, hi - start


def scatter_batch(full, src=0, group=None):
    """``full`` (B, ...) on ``src`` (None elsewhere) -> this rank's shard.  One metadata broadcast and ONE
    ``dist.scatter`` (NCCL on GPUs: grouped sends over NVLink; gloo in the CPU tests); the windows are views of
    ``full``, nothing is copied on the source."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = [None]
    if rank == src:
        meta = [(tuple(full.shape), full.dtype, full.device.type)]
    dist.broadcast_object_list(meta, src=src, group=group)
    shape, dtype, devtype = meta[0]
    if shape[0] < 1:
        raise ValueError("scatter_batch needs at least one mixture")
    device = torch.device("cuda", torch.cuda.current_device()) if devtype == "cuda" else torch.device("cpu")
    start, per, a, b = _window(shape[0], rank, world)
    is_c = dtype.is_complex
    recv = torch.empty((per,) + tuple(shape[1:]), dtype=dtype, device=device)
    rview = torch.view_as_real(recv) if is_c else recv
    chunks = None
    if rank == src:
        full = full.contiguous()
        fview = torch.view_as_real(full) if is_c else full
        chunks = []
        for r in range(world):
            s0, _, _, _ = _window(shape[0], r, world)
            chunks.append(fview[s0:s0 + per])
    dist.scatter(rview, chunks, src=src, group=group)
    return recv[a:b]


def gather_batch(shard, n_batch, dst=0, group=None):
    """Inverse of :func:`scatter_batch`: returns the (n_batch, ...) tensor on ``dst``, None elsewhere.  ONE
    ``dist.gather``; full-size shards land directly in the result, short ones go through a window buffer."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per = _per_rank(n_batch, world)
    is_c = shard.dtype.is_complex
    send = shard.contiguous()
    if send.shape[0] != per:  # short (or empty) shard: pad to the common window size
        pad = torch.zeros((per,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
        pad[:send.shape[0]].copy_(send)
        send = pad
    sview = torch.view_as_real(send) if is_c else send
    if rank != dst:
        dist.gather(sview, None, dst=dst, group=group)
        return None
    full = torch.empty((n_batch,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
    fview = torch.view_as_real(full) if is_c else full
    bufs, short = [], []
    for r in range(world):
        lo, hi = shard_range(n_batch, r, world)
        if hi - lo == per:
            bufs.append(fview[lo:hi])
        else:
            tmp = torch.empty_like(sview)
            bufs.append(tmp)
            short.append((lo, hi, tmp))
    dist.gather(sview, bufs, dst=dst, group=group)
    for lo, hi, tmp in short:
        fview[lo:hi].copy_(tmp[:hi - lo])
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
