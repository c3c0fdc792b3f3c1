exec(open(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'sens_whitened.py')).read().split("ref = run('')")[0])
# NMF operand split emulation: products (hi+lo)(hi+lo) minus lo*lo
def parts(x, mode):
    x = np.asarray(x, np.float32)
    hi = bf16_trunc(x) if mode == 'trunc' else bf16_rn(x)
    lo = bf16_rn(x - hi)
    return hi.astype(np.float64), lo.astype(np.float64)
def mm3(sub, A, B, mode):
    ah, al = parts(A, mode); bh, bl = parts(B, mode)
    return np.einsum(sub, ah, bh) + np.einsum(sub, ah, bl) + np.einsum(sub, al, bh)
def run_nmf(mode, which, n_iter=5, fp32_state=False):
    st = O.init_state(X, T0.copy(), V0.copy(), W=W0.copy(), spatial_algorithm='IP2')
    fl = S.max_flooring
    for _ in range(n_iter):
        P = np.abs(O._Y(st)) ** 2
        T, V = st['T'], st['V']
        R = mm3('nik,nkj->nij', T, V, mode) if 'R' in which else np.einsum('nik,nkj->nij', T, V)
        A = P / R ** 2; Bq = 1 / R
        if 'B' in which:
            num = mm3('nkj,nij->nik', V, A, mode); den = mm3('nkj,nij->nik', V, Bq, mode)
        else:
            num = np.einsum('nkj,nij->nik', V, A); den = np.einsum('nkj,nij->nik', V, Bq)
        st['T'] = fl(np.sqrt(num / den) * T)
        T = st['T']
        R = mm3('nik,nkj->nij', T, V, mode) if 'R' in which else np.einsum('nik,nkj->nij', T, V)
        A = P / R ** 2; Bq = 1 / R
        if 'A' in which:
            num = mm3('nik,nij->nkj', T, A, mode); den = mm3('nik,nij->nkj', T, Bq, mode)
        else:
            num = np.einsum('nik,nij->nkj', T, A); den = np.einsum('nik,nij->nkj', T, Bq)
        st['V'] = fl(np.sqrt(num / den) * V)
        if fp32_state:
            st['T'] = st['T'].astype(np.float32).astype(np.float64); st['V'] = st['V'].astype(np.float32).astype(np.float64)
        O.update_spatial(st, spatial_algorithm='IP2')
        O.normalize(st)
    Tn, Vn = st['T'].copy(), st['V'].copy()
    O.restore_scale(st)
    return st['Y'], Tn, Vn
ref, Tr, Vr = run_nmf('rn', '')
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
for mode in ['trunc', 'rn']:
    for which in ['R', 'B', 'A', 'RBA']:
        y, t, v = run_nmf(mode, which)
        print('%-6s %-4s relerr Y %.2e T %.2e V %.2e' % (mode, which, rel(y, ref), rel(t, Tr), rel(v, Vr)))
y, t, v = run_nmf('rn', '', fp32_state=True)
print('fp32 state T,V only: Y %.2e T %.2e V %.2e' % (rel(y, ref), rel(t, Tr), rel(v, Vr)))
