exec(open(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'sens_whitened.py')).read().split("def run(variant):")[0])
def run(phimode, gmode, n_iter=5):
    st = O.init_state(X, T0.copy(), V0.copy(), W=W0.copy(), spatial_algorithm='IP2')
    f = {'exact': lambda x: np.asarray(x, np.float64), 'f32': lambda x: np.asarray(x, np.float32).astype(np.float64),
         'rn2': lambda x: split(x, 'rn'), 'tr2': lambda x: split(x, 'trunc'), 'rn3': split3}
    orig_wc = S.weighted_covariance
    def wc(Xx, phi):
        Ni, Ii, Jj = Xx.shape
        U = np.empty((Ii, Ni, Ni, Ni), np.complex128)
        ph = f[phimode](phi)
        for a in range(Ni):
            for c in range(Ni):
                g = (Xx[a] * np.conj(Xx[c]))
                g = f[gmode](g.real.astype(np.float32)) + 1j * f[gmode](g.imag.astype(np.float32))
                U[:, :, a, c] = np.einsum('nij,ij->in', ph, g) / Jj
        return U
    S.weighted_covariance = wc
    try:
        for _ in range(n_iter):
            O.update_once(st, spatial_algorithm='IP2')
        O.restore_scale(st)
        return st['Y']
    finally:
        S.weighted_covariance = orig_wc
ref = run('exact', 'exact')
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
for pm, gm in [('f32', 'f32'), ('rn2', 'f32'), ('f32', 'rn2'), ('rn2', 'rn2'), ('rn2', 'rn3'), ('rn3', 'rn3'), ('rn3', 'rn2'), ('tr2','tr2')]:
    print('phi %-4s G %-4s relerr Y %.2e' % (pm, gm, rel(run(pm, gm), ref)))
