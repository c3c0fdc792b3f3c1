#!/usr/bin/env python
"""Where the end-to-end time of GaussILRMA.__call__(pinned host tensor) goes at config 2: PCIe copies alone and
together, the chunked pipeline on device-resident data, the host-side overhead (n_iter = 0), and the full call."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ssspy_b200.bss import GaussILRMA  # noqa: E402
from ssspy_b200.utils.synth import make_nmf_init  # noqa: E402

B, N, I, J, K, n_iter = 64, 2, 1025, 512, 16, 20
g = torch.Generator().manual_seed(0)
X = torch.complex(torch.randn(B, N, I, J, generator=g), torch.randn(B, N, I, J, generator=g)).pin_memory()
Yh = torch.empty_like(X).pin_memory()
Xd = torch.empty_like(X, device="cuda")
Yd = torch.empty_like(Xd)
T0, V0 = make_nmf_init(N, I, J, K)


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2]


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        Xd.copy_(X, non_blocking=True)
    with torch.cuda.stream(s2):
        Yh.copy_(Yd, non_blocking=True)


def call(x, n, chunk=None):
    m = GaussILRMA(n_basis=K, record_loss=False)
    if chunk:
        m.chunk_size = chunk
    return m(x, n_iter=n, basis=T0, activation=V0)


out = {"bytes_each_way_MB": X.numel() * 8 / 1e6,
       "h2d_ms": timed(lambda: Xd.copy_(X, non_blocking=True)),
       "d2h_ms": timed(lambda: Yh.copy_(Yd, non_blocking=True)),
       "h2d_and_d2h_together_ms": timed(both)}
call(X, 1)
out["call_host_tensor_n_iter_0_ms"] = timed(lambda: call(X, 0))
out["call_host_tensor_n_iter_20_ms"] = timed(lambda: call(X, n_iter))
out["call_device_tensor_n_iter_20_chunk8_ms"] = timed(lambda: call(Xd, n_iter, 8))
out["call_device_tensor_n_iter_20_default_ms"] = timed(lambda: call(Xd, n_iter))
out["call_device_tensor_n_iter_0_chunk8_ms"] = timed(lambda: call(Xd, 0, 8))
print(json.dumps(out))
