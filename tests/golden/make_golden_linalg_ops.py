#!/usr/bin/env python
"""Golden vectors of the standalone ssspy.linalg operators used by the IPA path: cbrt, solve_cubic, lqpqm2
(ssspy/linalg/cubic.py, polynomial.py, lqpqm.py).  Runs the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_linalg_ops.py      # -> tests/golden/linalg_ops.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.environ.get("SSSPY_REF", "/root/reference"))
from ssspy.linalg import cbrt, lqpqm2, solve_cubic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261017)
out = {}

# cbrt: real (both signs, zero) and complex
xr = np.concatenate([rng.standard_normal(40) * 10.0 ** rng.integers(-6, 6, 40), [0.0, -8.0, 27.0]])
xc = (rng.standard_normal(40) + 1j * rng.standard_normal(40)) * 10.0 ** rng.integers(-4, 4, 40)
xc = np.concatenate([xc, [-8.0 + 0j, 1j, -1j]])
out.update(cbrt_real_in=xr, cbrt_real_out=cbrt(xr), cbrt_cplx_in=xc, cbrt_cplx_out=cbrt(xc))

# solve_cubic: real monic (incl. P == 0: A = B = 0 and the shifted triple root), complex monic, general with D
A = rng.standard_normal(48)
B = rng.standard_normal(48)
C = rng.standard_normal(48)
A[:3], B[:3] = 0.0, 0.0                      # x^3 + C = 0: P == 0
A[3:6] = 3.0 * rng.standard_normal(3)
B[3:6] = A[3:6] ** 2 / 3.0                    # P == 0 with A != 0 (exactly representable? no: kept as a near-singular case)
out.update(cubic_A=A, cubic_B=B, cubic_C=C, cubic_roots=solve_cubic(A, B, C), cubic_first=solve_cubic(A, B, C, all=False))
Ac, Bc, Cc = (rng.standard_normal(32) + 1j * rng.standard_normal(32) for _ in range(3))
out.update(cubic_cA=Ac, cubic_cB=Bc, cubic_cC=Cc, cubic_croots=solve_cubic(Ac, Bc, Cc))
A4, B4, C4, D4 = (rng.standard_normal((4, 8)) for _ in range(4))
A4 = np.where(np.abs(A4) < 0.1, 0.5, A4)
out.update(cubic_gA=A4, cubic_gB=B4, cubic_gC=C4, cubic_gD=D4, cubic_groots=solve_cubic(A4, B4, C4, D4))

# lqpqm2: positive semidefinite H of size M = 1 .. 5, v != 0 (the v = 0 branch is a documented deviation)
for M in (1, 2, 3, 5):
    n = 24
    G = rng.standard_normal((n, M, M + 2)) + 1j * rng.standard_normal((n, M, M + 2))
    H = G @ np.conj(np.swapaxes(G, -2, -1)) / (M + 2)
    H[: n // 3] *= 1e-3
    v = rng.standard_normal((n, M)) + 1j * rng.standard_normal((n, M))
    v[n // 2:] *= 0.05
    z = rng.random(n) * 2.0
    z[::5] = 0.0
    for it in (1, 10):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            y = lqpqm2(H, v, z, max_iter=it)
        out["lqpqm2_M%d_it%d_out" % (M, it)] = y
    out.update({"lqpqm2_M%d_H" % M: H, "lqpqm2_M%d_v" % M: v, "lqpqm2_M%d_z" % M: z})

np.savez_compressed(os.path.join(HERE, "linalg_ops.npz"), **out)
print("wrote", os.path.join(HERE, "linalg_ops.npz"), len(out), "arrays")
