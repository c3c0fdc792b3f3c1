"""Differential test of the oracle against the UNMODIFIED reference imported live (build container only: skipped when
/root/reference is absent, e.g. on the GPU box; the committed fixtures of tests/golden/ cover that case).  Random
shapes and option combinations beyond the fixed golden cases, fp64 on both sides."""
import functools
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("SSSPY_REF", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ssspy")), reason="reference not available")

TOL = 1e-9


def _ref():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import ssspy.bss.ilrma as rilrma
    import ssspy.bss.iva as riva
    from ssspy.special.flooring import add_flooring, max_flooring
    from ssspy.utils.select_pair import combination_pair_selector, sequential_pair_selector
    return rilrma, riva, add_flooring, max_flooring, combination_pair_selector, sequential_pair_selector


def _relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def _draw(seed):
    rng = np.random.default_rng(seed)
    N = int(rng.integers(2, 5))
    I, J, K = int(rng.integers(5, 20)), int(rng.integers(24, 60)), int(rng.integers(2, 7))
    return rng, N, I, J, K


@pytest.mark.parametrize("seed", range(12))
def test_gauss_ilrma_random_options_oracle_equals_reference(seed):
    from oracle import ilrma as oilrma
    from oracle import spatial as ospatial
    from ssspy_b200.utils.synth import make_mixture, make_nmf_init
    rilrma, _, add_flooring, max_flooring, comb, seq = _ref()
    rng, N, I, J, K = _draw(1000 + seed)
    spatial = ["IP", "IP2", "ISS", "ISS2", "IPA"][seed % 5]
    source = "ME" if seed % 4 == 3 else "MM"
    domain = 2 if (source == "ME" or seed % 3) else 1
    flooring = "add" if seed % 6 == 5 else "max"
    normalization = [True, "projection_back", False][seed % 3]
    use_comb = spatial in ("IP2", "ISS2") and seed % 2 == 0
    X = make_mixture(N, I, J, seed=500 + seed, mode="mix")
    T, V = make_nmf_init(N, I, J, K, seed=600 + seed)
    n_iter = 3
    ref_floor = functools.partial(add_flooring if flooring == "add" else max_flooring, eps=1e-10)
    m = rilrma.GaussILRMA(n_basis=K, spatial_algorithm=spatial, source_algorithm=source, domain=domain,
                          flooring_fn=ref_floor, pair_selector=comb if use_comb else None,
                          normalization=normalization, scale_restoration=True, record_loss=True, reference_id=0,
                          rng=np.random.default_rng(0))
    Y = m(X, n_iter=n_iter, basis=T, activation=V)
    pairs = list((comb if use_comb else seq)(N))
    st = oilrma.run(X, T, V, n_iter, p=float(domain),
                    floor=ospatial.add_flooring if flooring == "add" else ospatial.max_flooring,
                    spatial_algorithm=spatial, source_algorithm=source, normalization=normalization, pairs=pairs,
                    reference_id=0, scale_restoration=True)
    assert _relerr(st["Y"], Y) < TOL
    assert _relerr(st["T"], m.basis) < TOL
    assert _relerr(st["V"], m.activation) < TOL
    np.testing.assert_allclose(st["loss"], m.loss, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("seed", range(10))
def test_aux_iva_random_options_oracle_equals_reference(seed):
    from oracle import iva as oiva
    from oracle import spatial as ospatial
    from ssspy_b200.utils.synth import make_mixture
    _, riva, add_flooring, max_flooring, comb, seq = _ref()
    rng, N, I, J, _ = _draw(2000 + seed)
    spatial = ["IP", "IP2", "ISS", "ISS2", "IPA"][seed % 5]
    model = "gauss" if seed % 2 else "laplace"
    use_comb = spatial in ("IP2", "ISS2") and seed % 4 < 2
    X = make_mixture(N, I, J, seed=700 + seed, mode="mix")
    cls = riva.AuxGaussIVA if model == "gauss" else riva.AuxLaplaceIVA
    sr = [True, "minimal_distortion_principle", False][seed % 3]
    ref_id = seed % N
    m = cls(spatial_algorithm=spatial, flooring_fn=functools.partial(max_flooring, eps=1e-10),
            pair_selector=comb if use_comb else None, scale_restoration=sr, record_loss=True, reference_id=ref_id)
    n_iter = 3
    Y = m(X, n_iter=n_iter)
    pairs = list((comb if use_comb else seq)(N))
    st = oiva.run(X, n_iter, floor=ospatial.max_flooring, spatial_algorithm=spatial, model=model, pairs=pairs,
                  reference_id=ref_id, scale_restoration=sr)
    assert _relerr(st["Y"], Y) < TOL
    np.testing.assert_allclose(st["loss"], m.loss, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("seed", range(8))
def test_t_ggd_and_partitioned_ilrma_oracle_equals_reference(seed):
    """TILRMA / GGDILRMA (ilrma.py:1992, :3337) and the partitioning function (latent variable) on random shapes."""
    from oracle import ilrma as oilrma
    from oracle import spatial as ospatial
    from ssspy_b200.utils.synth import make_mixture, make_nmf_init
    rilrma, _, _, max_flooring, _, seq = _ref()
    rng, N, I, J, K = _draw(3000 + seed)
    kind = ["t", "ggd", "gauss"][seed % 3]
    partitioning = seed % 2 == 1
    spatial = ["IP", "ISS", "IP2", "ISS2"][seed % 4]
    prm = float(rng.uniform(2.0, 40.0)) if kind == "t" else float(rng.uniform(0.6, 1.9))
    X = make_mixture(N, I, J, seed=800 + seed, mode="mix")
    T, V = make_nmf_init(N, I, J, K, seed=900 + seed)
    kwargs = dict(basis=T, activation=V)
    Z0 = None
    if partitioning:
        Z0 = rng.random((N, K)) + 0.1
        Z0 = Z0 / Z0.sum(axis=0)
        T, V = T[0].copy(), V[0].copy()
        kwargs = dict(basis=T, activation=V, latent=Z0)
    common = dict(n_basis=K, spatial_algorithm=spatial, source_algorithm="MM", domain=2,
                  flooring_fn=functools.partial(max_flooring, eps=1e-10), partitioning=partitioning,
                  normalization=True, scale_restoration=True, record_loss=True, reference_id=0,
                  rng=np.random.default_rng(0))
    if kind == "t":
        m = rilrma.TILRMA(dof=prm, **common)
    elif kind == "ggd":
        m = rilrma.GGDILRMA(beta=prm, **common)
    else:
        m = rilrma.GaussILRMA(**common)
    n_iter = 3
    Y = m(X, n_iter=n_iter, **kwargs)
    st = oilrma.run(X, T, V, n_iter, p=2.0, floor=ospatial.max_flooring, spatial_algorithm=spatial,
                    source_algorithm="MM", normalization=True, pairs=list(seq(N)), reference_id=0,
                    scale_restoration=True, dist=(kind, prm if kind != "gauss" else None), Z=Z0)
    assert _relerr(st["Y"], Y) < TOL
    assert _relerr(st["T"], m.basis) < TOL
    assert _relerr(st["V"], m.activation) < TOL
    if partitioning:
        assert _relerr(st["Z"], m.latent) < TOL
    np.testing.assert_allclose(st["loss"], m.loss, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("seed", range(6))
def test_fdica_and_mnmf_oracle_equals_reference(seed):
    """AuxLaplaceFDICA with permutation alignment (fdica.py:1527) and FastGaussMNMF with injected state (mnmf.py)."""
    from oracle import fdica as ofdica
    from oracle import mnmf as omnmf
    from oracle import spatial as ospatial
    from ssspy_b200.utils.synth import make_mixture
    _, _, _, max_flooring, comb, seq = _ref()
    from ssspy.bss.fdica import AuxLaplaceFDICA
    from ssspy.bss.mnmf import FastGaussMNMF
    rng, N, I, J, K = _draw(4000 + seed)
    J = J + 60  # few frames make the per-bin problems of FDICA ill-conditioned (see make_golden_fdica.py)
    alg = "IP2" if seed % 2 else "IP"
    use_comb = alg == "IP2" and seed % 4 == 1
    floor = functools.partial(max_flooring, eps=1e-10)
    pairs = list((comb if use_comb else seq)(N))
    X = make_mixture(N, I, J, seed=1100 + seed, mode="mix")
    m = AuxLaplaceFDICA(spatial_algorithm=alg, flooring_fn=floor, pair_selector=comb if use_comb else None,
                        permutation_alignment=True, scale_restoration=True, record_loss=True, reference_id=0)
    Y = m(X, n_iter=3)
    st = ofdica.run(X, 3, floor=ospatial.max_flooring, spatial_algorithm=alg, pairs=pairs, reference_id=0,
                    permutation_alignment=True, scale_restoration=True)
    assert _relerr(st["Y"], Y) < 1e-7
    np.testing.assert_allclose(st["loss"], m.loss, rtol=1e-8, atol=1e-8)

    T, V, D = rng.random((N, I, K)), rng.random((N, K, J)), rng.random((I, N, N))
    Q = np.eye(N)[None] + 0.3 * (rng.standard_normal((I, N, N)) + 1j * rng.standard_normal((I, N, N)))
    mm = FastGaussMNMF(n_basis=K, n_sources=N, diagonalizer_algorithm=alg, flooring_fn=floor,
                       pair_selector=comb if use_comb else None, normalization=True, record_loss=True,
                       reference_id=0, rng=np.random.default_rng(0))
    Ym = mm(X, n_iter=3, basis=T, activation=V, spatial=D, diagonalizer=Q)
    sm = omnmf.run(X, T, V, Q, D, 3, floor=ospatial.max_flooring, algorithm=alg, pairs=pairs, normalization=True,
                   reference_id=0)
    assert _relerr(sm["Y"], Ym) < 1e-7
    assert _relerr(sm["T"], mm.basis) < 1e-8
    np.testing.assert_allclose(sm["loss"], mm.loss, rtol=1e-8, atol=1e-8)


def test_lqpqm2_singular_branch_is_documented():
    """lqpqm2 (lqpqm.py:13-119): the regular branch equals the reference; in the v = 0 branch the reference returns
    scale * (last ROW of the eigenvector matrix, LAPACK phases), the oracle scale * (eigenvector of the largest
    eigenvalue).  Pin what is common to both -- the last component and the norm -- and that the oracle's result is
    an eigenvector (known deviation)."""
    _ref()
    from ssspy.linalg import lqpqm2 as ref_lqpqm2
    from oracle import spatial as ospatial
    from oracle.linalg import lqpqm2
    rng = np.random.default_rng(5)
    n, M = 6, 3
    A = rng.standard_normal((n, M, M)) + 1j * rng.standard_normal((n, M, M))
    H = A @ A.conj().transpose(0, 2, 1)
    v = rng.standard_normal((n, M)) + 1j * rng.standard_normal((n, M))
    v[:3] = 0
    z = rng.random(n) * 0.1
    yr = ref_lqpqm2(H, v.copy(), z)
    yo = lqpqm2(H, v.copy(), z, ospatial.max_flooring)
    assert _relerr(yo[3:], yr[3:]) < 1e-10
    np.testing.assert_allclose(yo[:3, -1], yr[:3, -1], rtol=1e-10)
    np.testing.assert_allclose(np.linalg.norm(yo[:3], axis=-1), np.linalg.norm(yr[:3], axis=-1), rtol=1e-10)
    lam = np.linalg.eigvalsh(H[:3])[:, -1]
    Hy = np.einsum("bij,bj->bi", H[:3], yo[:3])
    assert _relerr(Hy, lam[:, None] * yo[:3]) < 1e-10
    assert np.abs(yo[:3] - yr[:3]).max() > 1e-3  # the deviation is real


@pytest.mark.parametrize("case", ["zero_frames", "silent_channel", "tiny", "huge", "dup_channel", "zero_bin"])
@pytest.mark.parametrize("spatial", ["IP", "IP2", "ISS", "ISS2", "IPA"])
def test_degenerate_inputs_same_outcome_as_reference(case, spatial):
    """Degenerate inputs: the oracle must end like the reference -- the same exception type (LinAlgError for an
    exactly singular per-bin matrix), or the same finite result."""
    import warnings
    from oracle import ilrma as oilrma
    from oracle import spatial as ospatial
    from ssspy_b200.utils.synth import make_mixture, make_nmf_init
    rilrma, _, _, max_flooring, _, seq = _ref()
    N, I, J, K = 3, 9, 40, 4
    X = make_mixture(N, I, J, seed=1, mode="mix")
    T, V = make_nmf_init(N, I, J, K, seed=2)
    if case == "zero_frames":
        X[:, :, :5] = 0
    elif case == "silent_channel":
        X[1] = 0
    elif case == "tiny":
        X = X * 1e-7
    elif case == "huge":
        X = X * 1e6
    elif case == "dup_channel":
        X[2] = X[0]
    elif case == "zero_bin":
        X[:, 3, :] = 0

    def outcome(fn):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                return None, fn()
            except Exception as e:  # noqa: BLE001 - the exception type is what is compared
                return type(e).__name__, None

    m = rilrma.GaussILRMA(n_basis=K, spatial_algorithm=spatial, flooring_fn=functools.partial(max_flooring, eps=1e-10),
                          record_loss=True, reference_id=0, rng=np.random.default_rng(0))
    err_r, Y = outcome(lambda: m(X, n_iter=3, basis=T, activation=V))
    err_o, st = outcome(lambda: oilrma.run(X, T, V, 3, floor=ospatial.max_flooring, spatial_algorithm=spatial,
                                            pairs=list(seq(N)), reference_id=0))
    assert err_r == err_o
    if err_r is None:
        assert _relerr(st["Y"], Y) < 1e-9 and _relerr(st["T"], m.basis) < 1e-9
    else:
        assert err_r == "LinAlgError"


def test_constructor_validation_mirrors_reference():
    """Every combination of constructor options (valid, invalid and misspelt) ends the same way in the host classes
    of ssspy_b200.bss and in the reference's: no error, or the same exception type with the same message."""
    import itertools
    import warnings
    _ref()
    import ssspy.bss.fdica as rfd
    import ssspy.bss.ilrma as rilrma
    import ssspy.bss.iva as riva
    import ssspy.bss.mnmf as rmn
    from ssspy_b200.bss import fdica as mfd
    from ssspy_b200.bss import ilrma as milrma
    from ssspy_b200.bss import iva as miva
    from ssspy_b200.bss import mnmf as mmn

    def outcome(fn):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                fn()
                return None
            except Exception as e:  # noqa: BLE001
                return type(e).__name__ + ": " + str(e)

    spatial = ["IP", "IP1", "IP2", "ISS", "ISS1", "ISS2", "IPA", "XX"]
    restoration = [True, False, "projection_back", "minimal_distortion_principle", "MDP", "bogus"]
    grid = dict(spatial_algorithm=spatial, source_algorithm=["MM", "ME", "ZZ"], domain=[1, 1.5, 2],
                partitioning=[False, True], normalization=[True, False, "power", "projection_back", "bogus"],
                scale_restoration=restoration)
    n = 0
    for vals in itertools.product(*grid.values()):
        kw = dict(zip(grid, vals))
        for name, extra in (("GaussILRMA", {}), ("TILRMA", {"dof": 5.0}), ("GGDILRMA", {"beta": 1.3})):
            want = outcome(lambda: getattr(rilrma, name)(n_basis=3, **extra, **kw))
            got = outcome(lambda: getattr(milrma, name)(n_basis=3, **extra, **kw))
            assert got == want, (name, kw)
            n += 1
    for sa, sr in itertools.product(spatial, restoration):
        for name in ("AuxLaplaceIVA", "AuxGaussIVA"):
            assert outcome(lambda: getattr(miva, name)(spatial_algorithm=sa, scale_restoration=sr)) == \
                outcome(lambda: getattr(riva, name)(spatial_algorithm=sa, scale_restoration=sr)), (name, sa, sr)
        assert outcome(lambda: mfd.AuxLaplaceFDICA(spatial_algorithm=sa, scale_restoration=sr)) == \
            outcome(lambda: rfd.AuxLaplaceFDICA(spatial_algorithm=sa, scale_restoration=sr)), (sa, sr)
        assert outcome(lambda: mmn.FastGaussMNMF(n_basis=3, diagonalizer_algorithm=sa)) == \
            outcome(lambda: rmn.FastGaussMNMF(n_basis=3, diagonalizer_algorithm=sa)), sa
    assert n == 12960


def test_repr_mirrors_reference():
    _ref()
    import ssspy.bss.fdica as rfd
    import ssspy.bss.ilrma as rilrma
    import ssspy.bss.iva as riva
    import ssspy.bss.mnmf as rmn
    from ssspy_b200.bss import fdica as mfd
    from ssspy_b200.bss import ilrma as milrma
    from ssspy_b200.bss import iva as miva
    from ssspy_b200.bss import mnmf as mmn
    cases = [("GaussILRMA", rilrma, milrma, dict(n_basis=4)),
             ("GaussILRMA", rilrma, milrma, dict(n_basis=4, spatial_algorithm="ISS2", partitioning=True)),
             ("GaussILRMA", rilrma, milrma, dict(n_basis=2, spatial_algorithm="IPA", normalization=False,
                                                  scale_restoration="MDP", record_loss=False, reference_id=1)),
             ("TILRMA", rilrma, milrma, dict(n_basis=4, dof=100)), ("GGDILRMA", rilrma, milrma, dict(n_basis=4, beta=1.5)),
             ("AuxLaplaceIVA", riva, miva, {}), ("AuxGaussIVA", riva, miva, dict(spatial_algorithm="IPA")),
             ("AuxIVA", riva, miva, dict(contrast_fn=None, d_contrast_fn=None)),
             ("AuxLaplaceFDICA", rfd, mfd, {}), ("AuxLaplaceFDICA", rfd, mfd, dict(spatial_algorithm="IP2")),
             ("FastGaussMNMF", rmn, mmn, dict(n_basis=3)), ("FastGaussMNMF", rmn, mmn, dict(n_basis=3, n_sources=2))]
    for name, ref_mod, new_mod, kw in cases:
        assert repr(getattr(new_mod, name)(**kw)) == repr(getattr(ref_mod, name)(**kw)), (name, kw)


def test_public_method_surface_covers_reference():
    """Every public attribute of the reference's separator classes exists on the host classes (drop-in surface)."""
    _ref()
    import ssspy.bss.fdica as rfd
    import ssspy.bss.ilrma as rilrma
    import ssspy.bss.iva as riva
    import ssspy.bss.mnmf as rmn
    from ssspy_b200.bss import fdica as mfd
    from ssspy_b200.bss import ilrma as milrma
    from ssspy_b200.bss import iva as miva
    from ssspy_b200.bss import mnmf as mmn
    for name, ref_mod, new_mod in (("GaussILRMA", rilrma, milrma), ("TILRMA", rilrma, milrma),
                                   ("GGDILRMA", rilrma, milrma), ("AuxIVA", riva, miva),
                                   ("AuxLaplaceIVA", riva, miva), ("AuxGaussIVA", riva, miva),
                                   ("AuxLaplaceFDICA", rfd, mfd), ("FastGaussMNMF", rmn, mmn)):
        want = {n for n in dir(getattr(ref_mod, name)) if not n.startswith("_")}
        have = {n for n in dir(getattr(new_mod, name)) if not n.startswith("_")}
        assert sorted(want - have) == [], name
