import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ilrma as oilrma
from ssspy_b200.bss import GaussILRMA
from ssspy_b200.utils.synth import make_mixture, make_nmf_init

def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))

N, I, J, K = 3, 33, 48, 5
X = make_mixture(N, I, J, seed=31, mode="mix")
T, V = make_nmf_init(N, I, J, K, seed=32)
for source in ("MM", "ME"):
    whole = GaussILRMA(n_basis=K, source_algorithm=source)
    whole(X, n_iter=0, basis=T, activation=V)
    print(source, "after call n_iter=0: T vs T0", relerr(whole.basis, T), "V vs V0", relerr(whole.activation, V),
          "W-I", np.abs(whole.demix_filter - np.eye(N)).max())
    whole.update_source_model()
    st = oilrma.init_state(X, T, V, None, "IP", None)
    oilrma.update_basis(st, source_algorithm=source)
    oilrma.update_activation(st, source_algorithm=source)
    print(source, "whole vs oracle: T", relerr(whole.basis, st["T"]), "V", relerr(whole.activation, st["V"]))
    parts = GaussILRMA(n_basis=K, source_algorithm=source)
    parts(X, n_iter=0, basis=T, activation=V)
    getattr(parts, "update_basis_" + source.lower())()
    st2 = oilrma.init_state(X, T, V, None, "IP", None)
    oilrma.update_basis(st2, source_algorithm=source)
    print(source, "parts basis vs oracle", relerr(parts.basis, st2["T"]), "vs whole", relerr(parts.basis, whole.basis),
          "vs T0", relerr(parts.basis, T))
    getattr(parts, "update_activation_" + source.lower())()
    print(source, "parts act vs whole", relerr(parts.activation, whole.activation))
