#!/usr/bin/env python
"""Times one update_once step of every BASELINE.json configuration that fits one GPU (device-resident
inputs generated on the GPU, record_loss=False, CUDA events, 3 warm-up + `--steps` timed steps) and
prints one JSON line per configuration with the algorithmic-bytes roofline fraction.
Not the driver's bench (that is bench.py); this is the per-row measurement table for DESIGN.md."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ssspy_b200 import _lib  # noqa: E402
from ssspy_b200.bss import AuxGaussIVA, AuxLaplaceIVA, FastGaussMNMF, GaussILRMA  # noqa: E402


def bytes_ilrma(N, I, J, K):
    return 8 * N * I * J + 2 * 4 * (N * I * K + N * K * J) + 2 * 8 * N * N * I


def bytes_iva_ip(N, I, J):
    return 8 * N * I * J + 2 * 8 * N * N * I + 2 * 4 * N * J


def bytes_iva_iss(N, I, J):
    return 2 * 8 * N * I * J + 2 * 4 * N * J


def run(name, make, X, abytes, steps, peak, **state):
    # per-kernel breakdown from a single plan on one stream (event deltas mean nothing across the chunk streams the
    # engine uses by default); the step time below is measured with the engine's default layout
    prof = make()
    prof.chunk_size = X.shape[0]
    prof(X, n_iter=0, **state)
    for _ in range(2):
        prof.update_once()
    torch.cuda.synchronize()
    _lib.call("ssb_profile_begin", torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        prof.update_once()
    kern = _lib.profile_end()
    del prof
    torch.cuda.empty_cache()
    sep = make()
    sep(X, n_iter=0, **state)
    for _ in range(3):
        sep.update_once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sep.update_once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    extra = {}
    if isinstance(sep, FastGaussMNMF):  # the per-(bin, frame) eigendecomposition path of config 5
        e0.record()
        sep._plan_call("ssb_plan_separate")
        e1.record()
        torch.cuda.synchronize()
        extra["separate_ms (Wiener filter, I*J Hermitian eigh per mixture)"] = round(e0.elapsed_time(e1), 3)
        extra["eigh_per_sec"] = round(X.shape[0] * X.shape[2] * X.shape[3] / (e0.elapsed_time(e1) * 1e-3), 0)
    B = X.shape[0]
    gbs = abytes * B / (ms * 1e-3) / 1e9
    print(json.dumps({"config": name, "batch": B, "ms_per_step": round(ms, 4),
                      "mixture_iterations_per_sec": round(B / (ms * 1e-3), 1),
                      "algorithmic_MB_per_step": round(abytes * B / 1e6, 1), "achieved_GBps": round(gbs, 1),
                      "hbm_frac": round(gbs / peak, 4),
                      "kernels_ms_per_step": {k[0]: round(k[2] / 2, 4) for k in kern}, **extra}), flush=True)
    del sep
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    g = torch.Generator(device="cuda").manual_seed(0)

    def randc(*shape):
        return torch.complex(torch.randn(*shape, device="cuda", generator=g), torch.randn(*shape, device="cuda", generator=g))

    def nmf(B, N, I, J, K):
        return dict(basis=torch.rand(B, N, I, K, device="cuda", generator=g) + 0.05,
                    activation=torch.rand(B, N, K, J, device="cuda", generator=g) + 0.05)

    cfgs = []
    cfgs.append(("c1 AuxLaplaceIVA-IP N=2 I=257 J=128 B=1", lambda: AuxLaplaceIVA("IP", record_loss=False, scale_restoration=False),
                 (1, 2, 257, 128), None, bytes_iva_ip(2, 257, 128)))
    for N in (2, 4, 8):
        cfgs.append(("c2 GaussILRMA-IP N=%d I=1025 J=512 K=16 B=64" % N,
                     lambda: GaussILRMA(16, "IP", record_loss=False, scale_restoration=False), (64, N, 1025, 512), 16,
                     bytes_ilrma(N, 1025, 512, 16)))
    cfgs.append(("c2b GaussILRMA-IP2 N=4 I=1025 J=512 K=16 B=64", lambda: GaussILRMA(16, "IP2", record_loss=False, scale_restoration=False),
                 (64, 4, 1025, 512), 16, bytes_ilrma(4, 1025, 512, 16)))
    cfgs.append(("c2c GaussILRMA-ISS N=4 I=1025 J=512 K=16 B=64", lambda: GaussILRMA(16, "ISS", record_loss=False, scale_restoration=False),
                 (64, 4, 1025, 512), 16, bytes_ilrma(4, 1025, 512, 16) + 8 * 4 * 1025 * 512))
    cfgs.append(("c3 AuxLaplaceIVA-ISS N=4 I=1025 J=512 B=256", lambda: AuxLaplaceIVA("ISS", record_loss=False, scale_restoration=False),
                 (256, 4, 1025, 512), None, bytes_iva_iss(4, 1025, 512)))
    cfgs.append(("c3b AuxLaplaceIVA-IP N=4 I=1025 J=512 B=256", lambda: AuxLaplaceIVA("IP", record_loss=False, scale_restoration=False),
                 (256, 4, 1025, 512), None, bytes_iva_ip(4, 1025, 512)))
    cfgs.append(("c3c AuxGaussIVA-IP2 N=4 I=1025 J=512 B=256", lambda: AuxGaussIVA("IP2", record_loss=False, scale_restoration=False),
                 (256, 4, 1025, 512), None, bytes_iva_ip(4, 1025, 512)))
    cfgs.append(("c4 GaussILRMA-IP2 N=8 I=2049 J=1024 K=32 B=64 (one GPU's shard of B=512)",
                 lambda: GaussILRMA(32, "IP2", record_loss=False, scale_restoration=False), (64, 8, 2049, 1024), 32,
                 bytes_ilrma(8, 2049, 1024, 32)))
    cfgs.append(("c5 FastGaussMNMF-IP N=4 I=1025 J=512 K=16 B=256", lambda: FastGaussMNMF(16, diagonalizer_algorithm="IP", record_loss=False),
                 (256, 4, 1025, 512), 16, bytes_ilrma(4, 1025, 512, 16) + 2 * 4 * 1025 * 4 * 4))
    cfgs.append(("c5b FastGaussMNMF-IP2 N=4 I=1025 J=512 K=16 B=256", lambda: FastGaussMNMF(16, diagonalizer_algorithm="IP2", record_loss=False),
                 (256, 4, 1025, 512), 16, bytes_ilrma(4, 1025, 512, 16) + 2 * 4 * 1025 * 4 * 4))
    for name, make, shape, K, ab in cfgs:
        if args.only and args.only not in name:
            continue
        B, N, I, J = shape
        X = randc(B, N, I, J)
        state = nmf(B, N, I, J, K) if K else {}
        run(name, make, X, ab, args.steps, peak, **state)
        del X, state
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
