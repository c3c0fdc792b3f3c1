"""IP2 error growth on the device: final Y against the fp64 oracle after 2 / 5 / 12 iterations, fused and modular
kernels (run on the GPU box; informs the tolerance discussion in DESIGN.md 'IP2 sensitivity')."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import ilrma as oilrma  # noqa: E402
from ssspy_b200.bss import GaussILRMA  # noqa: E402
from ssspy_b200.utils.synth import make_mixture, make_nmf_init  # noqa: E402


def relerr(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


for (N, I, J, K) in [(2, 1025, 512, 16), (4, 257, 512, 16), (8, 129, 1024, 32)]:
    X = make_mixture(N, I, J, seed=2000 + N, mode="mix")
    T, V = make_nmf_init(N, I, J, K)
    for n_iter in (2, 5, 12):
        st = oilrma.run(X, T, V, n_iter, spatial_algorithm="IP2")
        row = {"N": N, "I": I, "J": J, "K": K, "n_iter": n_iter}
        for fast in (True, False):
            m = GaussILRMA(n_basis=K, spatial_algorithm="IP2")
            m.fast_path = fast
            Y = m(X, n_iter=n_iter, basis=T, activation=V)
            key = "fused" if fast else "modular"
            row[key + "_Y"] = relerr(Y, st["Y"])
            row[key + "_T"] = relerr(m.basis, st["T"])
            row[key + "_loss"] = float(np.max(np.abs(np.asarray(m.loss) / np.asarray(st["loss"]) - 1)))
        print(json.dumps(row), flush=True)
