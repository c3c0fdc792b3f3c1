#!/usr/bin/env python
"""Coarse host-side timeline of GaussILRMA.__call__ on a pinned host tensor (config 2): where the end-to-end time
goes besides the PCIe copies."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ssspy_b200.bss import GaussILRMA  # noqa: E402
from ssspy_b200.bss import _engine  # noqa: E402

B, N, I, J, K, n_iter = 64, 2, 1025, 512, 16, 20
g = torch.Generator().manual_seed(0)
X = torch.complex(torch.randn(B, N, I, J, generator=g), torch.randn(B, N, I, J, generator=g)).pin_memory()
rng = np.random.default_rng(0)
T0 = rng.random((B, N, I, K)) + 0.05
V0 = rng.random((B, N, K, J)) + 0.05


def stamp(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("  %-28s %7.2f ms (synchronised)" % (label, (t1 - t0) * 1e3))
    return time.perf_counter()


for rep in range(3):
    print("rep", rep)
    m = GaussILRMA(n_basis=K, spatial_algorithm="IP", record_loss=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.input = X
    t = stamp("input setter", t0)
    m._reset(flooring_fn=m.flooring_fn, basis=T0, activation=V0)
    t = stamp("_reset (state upload, W X)", t)
    m._ensure_plan()
    t = stamp("_ensure_plan (+ H2D of X)", t)
    m._stock_pipeline(n_iter, True, pb=True)
    t = stamp("_stock_pipeline (+ D2H)", t)
    print("  total %.2f ms" % ((t - t0) * 1e3))
    t0 = time.perf_counter()
    Y = GaussILRMA(n_basis=K, spatial_algorithm="IP", record_loss=False)(X, n_iter=n_iter, basis=T0, activation=V0)
    torch.cuda.synchronize()
    print("  plain __call__ %.2f ms" % ((time.perf_counter() - t0) * 1e3))
