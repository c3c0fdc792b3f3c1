#!/usr/bin/env python
"""Host <-> device copy bandwidth of every rank alone and of all ranks at once (torchrun, one rank per GPU): explains
how the end-to-end (host tensor in / out) numbers of bench.py scale with the number of GPUs that share one host.
Rank 0 prints one JSON line."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 64 * 2 * 1025 * 512  # config 2 shard: 537 MB each way
h, h2 = torch.empty(n, dtype=torch.complex64).pin_memory(), torch.empty(n, dtype=torch.complex64).pin_memory()
d, d2 = torch.empty(n, dtype=torch.complex64, device="cuda"), torch.empty(n, dtype=torch.complex64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 8 / 1e9


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def t(fn, rep=3):
    best = 1e9
    for _ in range(rep):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def solo(fn):
    """every rank in turn, the others idle"""
    out = 0.0
    for r in range(world):
        if r == rank:
            out = t2(fn)
        if world > 1:
            dist.barrier()
    return out


def t2(fn, rep=3):
    best = 1e9
    for _ in range(rep):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


res = {"solo_h2d_gbs": gb / solo(lambda: d.copy_(h, non_blocking=True)),
       "solo_d2h_gbs": gb / solo(lambda: h2.copy_(d2, non_blocking=True)),
       "all_h2d_gbs": gb / t(lambda: d.copy_(h, non_blocking=True)),
       "all_d2h_gbs": gb / t(lambda: h2.copy_(d2, non_blocking=True)),
       "all_both_gbs_each_way": gb / t(both)}
if world > 1:
    vals = torch.tensor([res[k] for k in sorted(res)], device="cuda")
    gathered = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(gathered, vals)
    if rank == 0:
        keys = sorted(res)
        out = {k: [round(float(g[i]), 1) for g in gathered] for i, k in enumerate(keys)}
        out["aggregate_all_h2d_gbs"] = round(sum(out["all_h2d_gbs"]), 1)
        out["aggregate_all_both_gbs_each_way"] = round(sum(out["all_both_gbs_each_way"]), 1)
        out["world"] = world
        out["cpu_affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))]
        print(json.dumps(out))
    dist.destroy_process_group()
else:
    print(json.dumps({k: round(v, 1) for k, v in res.items()}))
