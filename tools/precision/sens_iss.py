import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import ilrma as O
from ssspy_b200.utils.synth import make_batch, make_nmf_init
N, I, J, K = 4, 1025, 512, 16
X = make_batch(1, N, I, J, config_id=2, mode="mix")[0]
T, V = make_nmf_init(N, I, J, K, seed=42)
ref = O.run(X, T, V, 2, spatial_algorithm="ISS")["Y"]
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
rng = np.random.default_rng(0)
for eps in [1e-7, 1e-6]:
    Xp = X * (1 + eps * rng.standard_normal(X.shape))
    y = O.run(Xp, T, V, 2, spatial_algorithm="ISS")["Y"]
    e = np.linalg.norm(y - ref, axis=(0, 2)) / np.linalg.norm(ref, axis=(0, 2))
    print("X noise %.0e: relerr Y %.2e; worst bins %s %s" % (eps, rel(y, ref), np.argsort(e)[-3:], np.sort(e)[-3:]))
    Tp = T * (1 + eps * rng.standard_normal(T.shape))
    y = O.run(X, Tp, V, 2, spatial_algorithm="ISS")["Y"]
    e = np.linalg.norm(y - ref, axis=(0, 2)) / np.linalg.norm(ref, axis=(0, 2))
    print("T noise %.0e: relerr Y %.2e; worst bins %s %s" % (eps, rel(y, ref), np.argsort(e)[-3:], np.sort(e)[-3:]))
X32 = X.astype(np.complex64).astype(np.complex128)
y = O.run(X32, T, V, 2, spatial_algorithm="ISS")["Y"]
e = np.linalg.norm(y - ref, axis=(0, 2)) / np.linalg.norm(ref, axis=(0, 2))
print("X complex64: relerr Y %.2e; worst bins %s %s" % (rel(y, ref), np.argsort(e)[-3:], np.sort(e)[-3:]))
