#!/usr/bin/env python
"""End-to-end time of GaussILRMA.__call__(pinned host tensor) at BASELINE config 2 for several chunk layouts of the
host-tensor pipeline (SSB_HOST_LAYOUT / separator.host_layout).  One JSON line per layout."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ssspy_b200.bss import GaussILRMA  # noqa: E402
from ssspy_b200.utils.synth import make_nmf_init  # noqa: E402

B, N, I, J, K, n_iter = 64, 2, 1025, 512, 16, 20
g = torch.Generator().manual_seed(0)
X = torch.complex(torch.randn(B, N, I, J, generator=g), torch.randn(B, N, I, J, generator=g)).pin_memory()
T0, V0 = make_nmf_init(N, I, J, K)
layouts = [None, "16,16,16,8,4,2,2", "16,16,16,16", "12,12,12,12,8,4,2,2", "8,8,8,8,8,8,8,4,2,2", "4,12,16,16,8,4,2,2",
           "32,16,8,4,2,2", "16,16,12,8,6,4,2"]
for streams in (4, 8):
    for lay in layouts:
        ts = []
        for rep in range(4):
            m = GaussILRMA(n_basis=K, spatial_algorithm="IP", record_loss=False)  # a new separator per call, as bench.py
            m.host_layout, m.n_streams = lay, streams
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            Y = m(X, n_iter=n_iter, basis=T0, activation=V0)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print(json.dumps({"layout": lay, "streams": streams, "ms": [round(t * 1e3, 2) for t in ts],
                          "mixture_iterations_per_sec": round(B * n_iter / min(ts[1:]), 1)}), flush=True)
