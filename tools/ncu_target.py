#!/usr/bin/env python
"""Small launch sequence for ncu at BASELINE config 2 (or --sources N): one plan over the whole batch, a few steps of
ssb_run (the engine path bench.py times when --chunked is given: four chunk plans on four streams)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

sys.argv = [sys.argv[0]] + sys.argv[1:]
ap = argparse.ArgumentParser()
ap.add_argument("--sources", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--chunked", action="store_true")
ap.add_argument("--config", type=int, default=2)
args = ap.parse_args()
import bench  # noqa: E402
from ssspy_b200 import bss  # noqa: E402
from ssspy_b200.utils.synth import make_nmf_init  # noqa: E402

wl = dict(bench.CONFIGS[args.config])
if args.sources:
    wl["n_sources"] = args.sources if args.config == 2 else wl["n_sources"]
N, I, J, K, B = wl["n_sources"], wl["n_bins"], wl["n_frames"], wl["n_basis"], wl["batch"]
Xd = bench.synth_batch_device(B, N, I, J, seed=1000 * args.config, torch=torch)
if wl["cls"] == "GaussILRMA":
    T0, V0 = make_nmf_init(N, I, J, K, seed=42)
    m = bss.GaussILRMA(n_basis=K, spatial_algorithm=wl["spatial"], record_loss=False, scale_restoration=False)
    st = dict(basis=T0, activation=V0)
elif wl["cls"] == "AuxLaplaceIVA":
    m = bss.AuxLaplaceIVA(spatial_algorithm=wl["spatial"], record_loss=False, scale_restoration=False)
    st = {}
else:
    import numpy as np
    m = bss.FastGaussMNMF(n_basis=K, diagonalizer_algorithm=wl["spatial"], record_loss=False, rng=np.random.default_rng(7))
    st = {}
if not args.chunked:
    m.chunk_size = B
m(Xd, n_iter=0, **st)
if hasattr(m, "run_iterations"):
    m.run_iterations(args.steps)
else:
    for _ in range(args.steps):
        m.update_once()
torch.cuda.synchronize()
print("done")
