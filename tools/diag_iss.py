"""Per-bin error of GaussILRMA-ISS N=4 at I=1025, J=512 after 2 iterations (fast path and modular kernels) against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ilrma as oilrma
from ssspy_b200.bss import GaussILRMA
from ssspy_b200.utils.synth import make_batch, make_nmf_init
N, I, J, K = 4, 1025, 512, 16
X = make_batch(2, N, I, J, config_id=2, mode="mix")
T, V = make_nmf_init(N, I, J, K, seed=42)
for n_iter in (1, 2, 3, 5):
    st = oilrma.run(X[0], T, V, n_iter, spatial_algorithm="ISS")
    for fast in (True, False):
        m = GaussILRMA(n_basis=K, spatial_algorithm="ISS")
        m.fast_path = fast
        Y = m(X[0], n_iter=n_iter, basis=T, activation=V)
        e = np.linalg.norm(Y - st["Y"], axis=(0, 2)) / np.linalg.norm(st["Y"], axis=(0, 2))
        tot = np.linalg.norm(Y - st["Y"]) / np.linalg.norm(st["Y"])
        o = np.argsort(e)[-4:]
        print("n_iter %d fast=%s: relerr Y %.2e T %.2e; worst bins %s %s; median bin %.2e" % (
            n_iter, fast, tot, np.linalg.norm(m.basis - st["T"]) / np.linalg.norm(st["T"]), o, e[o], np.median(e)))
