#!/usr/bin/env python
"""End-to-end (pinned host tensor in -> pinned host tensor out) time of GaussILRMA.__call__ at BASELINE config 2 for
several chunk / stream layouts of the engine (ssspy_b200/bss/_engine.py).  Prints one JSON line per layout."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ssspy_b200.bss import GaussILRMA  # noqa: E402

B, N, I, J, K, n_iter = 64, 2, 1025, 512, 16, 20
g = torch.Generator().manual_seed(0)
X = torch.complex(torch.randn(B, N, I, J, generator=g), torch.randn(B, N, I, J, generator=g)).pin_memory()
rng = np.random.default_rng(0)
T0 = rng.random((B, N, I, K)) + 0.05
V0 = rng.random((B, N, K, J)) + 0.05
for chunk, streams in ((None, 3), (16, 3), (8, 3), (8, 4), (4, 4), (8, 6), (4, 6), (2, 6)):
    ts = []
    for rep in range(3):
        m = GaussILRMA(n_basis=K, spatial_algorithm="IP", record_loss=False)
        m.chunk_size, m.n_streams = chunk, streams
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Y = m(X, n_iter=n_iter, basis=T0, activation=V0)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"chunk": chunk, "streams": streams, "ms": [round(t * 1e3, 2) for t in ts],
                      "mixture_iterations_per_sec": round(B * n_iter / min(ts[1:]), 1)}), flush=True)
