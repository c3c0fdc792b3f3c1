import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import ilrma as O, spatial as S
from ssspy_b200.utils.synth import make_mixture, make_nmf_init

def bf16_trunc(x):
    u = np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xffff0000)
    return u.view(np.float32)
def bf16_rn(x):
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7fff + ((u >> 16) & 1)) & 0xffff0000).astype(np.uint32)
    return r.view(np.float32)
def split(x, mode):
    x = np.asarray(x, np.float32)
    hi = bf16_trunc(x) if mode == 'trunc' else bf16_rn(x)
    lo = bf16_rn(x - hi)
    return hi.astype(np.float64) + lo.astype(np.float64)
def split3(x):
    x = np.asarray(x, np.float32)
    hi = bf16_rn(x); r = x - hi; mid = bf16_rn(r); lo = bf16_rn(r - mid)
    return hi.astype(np.float64) + mid.astype(np.float64) + lo.astype(np.float64)

N, I, J, K, n_iter = 8, 48, 1024, 32, 5
X = make_mixture(N, I, J, seed=48000, mode='mix')
X = X.astype(np.complex64).astype(np.complex128)
T0, V0 = make_nmf_init(N, I, J, K, seed=42)
# whitened domain: Z = L^-1 X, W~0 = L  (C = X X^H / J = L L^H)
Xi = X.transpose(1, 0, 2)
C = Xi @ np.conj(Xi.transpose(0, 2, 1)) / J
L = np.linalg.cholesky(C)
Z = np.linalg.solve(L, Xi).transpose(1, 0, 2)
Z = Z.astype(np.complex64).astype(np.complex128)
X = Z
W0 = L.copy()

def run(variant):
    st = O.init_state(X, T0.copy(), V0.copy(), W=W0.copy(), spatial_algorithm='IP2')
    orig_wc = S.weighted_covariance
    orig_rec = O.reconstruct
    def wc(Xx, phi):
        if 'Gsplit' in variant:
            # emulate split of phi and G (rn, 2^-18), exact accumulate
            Ni, Ii, Jj = Xx.shape
            U = np.empty((Ii, Ni, Ni, Ni), np.complex128)
            ph = split(phi, 'rn')
            for a in range(Ni):
                for c in range(Ni):
                    g = (Xx[a] * np.conj(Xx[c]))
                    g = split(g.real.astype(np.float32), 'rn') + 1j * split(g.imag.astype(np.float32), 'rn')
                    U[:, :, a, c] = np.einsum('nij,ij->in', ph, g) / Jj
            return U
        U = orig_wc(Xx, phi)
        if 'U32' in variant:
            U = U.astype(np.complex64).astype(np.complex128)
        return U
    def rec(stt):
        if 'Rsplit_trunc' in variant and getattr(rec, 'in_spatial', False):
            return np.einsum('nik,nkj->nij', split(stt['T'], 'trunc'), split(stt['V'], 'trunc'))
        if 'Rsplit_rn' in variant and getattr(rec, 'in_spatial', False):
            return np.einsum('nik,nkj->nij', split(stt['T'], 'rn'), split(stt['V'], 'rn'))
        if 'R32' in variant and getattr(rec, 'in_spatial', False):
            return orig_rec(stt).astype(np.float32).astype(np.float64)
        return orig_rec(stt)
    S.weighted_covariance = wc
    O.reconstruct = rec
    orig_us = O.update_spatial
    def us(stt, *a, **k):
        rec.in_spatial = True
        try:
            return orig_us(stt, *a, **k)
        finally:
            rec.in_spatial = False
    O.update_spatial = us
    try:
        for _ in range(n_iter):
            O.update_once(st, spatial_algorithm='IP2')
            if 'W32' in variant:
                st['W'] = st['W'].astype(np.complex64).astype(np.complex128)
        O.restore_scale(st)
        return st['Y']
    finally:
        S.weighted_covariance = orig_wc; O.reconstruct = orig_rec; O.update_spatial = orig_us

ref = run('')
for v in ['U32', 'R32', 'Rsplit_trunc', 'Rsplit_rn', 'Gsplit', 'W32', 'Rsplit_trunc+Gsplit+U32+W32', 'Rsplit_rn+Gsplit+U32+W32']:
    y = run(v)
    print('%-32s relerr Y %.2e' % (v, np.linalg.norm(y - ref) / np.linalg.norm(ref)))
