"""Pin the oracle (oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import ilrma as oilrma
from oracle import iva as oiva
from oracle import linalg as olinalg
from oracle import spatial as ospatial
from oracle.projection_back import projection_back

from helpers import FLOORS, dist_arg, golden_cases, ipa_arg, load, norm_arg, phase_align_rows, relerr, sr_arg

TOL = 1e-9


@pytest.mark.parametrize("name", golden_cases("ilrma_") + golden_cases("tilrma_") + golden_cases("ggdilrma_"))
def test_ilrma_oracle_matches_reference(name):
    g = load(name)
    ref_id = None if int(g["reference_id"]) < 0 else int(g["reference_id"])
    st = oilrma.run(g["X"], g["T0"], g["V0"], int(g["n_iter"]), W=g.get("W0"), p=float(g["domain"]),
                    floor=FLOORS[str(g["flooring"])], spatial_algorithm=str(g["spatial"]),
                    source_algorithm=str(g["source"]), normalization=norm_arg(g["normalization"]),
                    pairs=[tuple(p) for p in g["pairs"]], reference_id=ref_id,
                    scale_restoration=sr_arg(g["scale_restoration"]), snapshots=True, dist=dist_arg(g),
                    Z=g.get("Z0"), ipa=ipa_arg(g))
    if "Z" in g:
        assert relerr(st["Z"], g["Z"]) < TOL
    assert relerr(st["Y"], g["Y"]) < TOL
    assert relerr(st["T"], g["T"]) < TOL
    assert relerr(st["V"], g["V"]) < TOL
    assert relerr(st["snapshots"][0]["T"], g["T_first_iter"]) < TOL
    assert relerr(st["snapshots"][0]["V"], g["V_first_iter"]) < TOL
    np.testing.assert_allclose(st["loss"], g["loss"], rtol=1e-10, atol=1e-9)
    if "W" in g:
        assert relerr(st["W"], g["W"]) < TOL
        W1 = st["snapshots"][0]["W"]
        if str(g["spatial"]) == "IP2":  # eigenvector phase is LAPACK's choice
            W1 = phase_align_rows(W1, g["W_first_iter"])
        assert relerr(W1, g["W_first_iter"]) < TOL


@pytest.mark.parametrize("name", golden_cases("iva_"))
def test_iva_oracle_matches_reference(name):
    g = load(name)
    st = oiva.run(g["X"], int(g["n_iter"]), W=g.get("W0"), floor=FLOORS[str(g["flooring"])],
                  spatial_algorithm=str(g["spatial"]), model=str(g["model"]),
                  pairs=[tuple(p) for p in g["pairs"]], reference_id=int(g["reference_id"]),
                  scale_restoration=sr_arg(g["scale_restoration"]), ipa=ipa_arg(g))
    assert relerr(st["Y"], g["Y"]) < TOL
    np.testing.assert_allclose(st["loss"], g["loss"], rtol=1e-10, atol=1e-9)
    if "W" in g:
        assert relerr(st["W"], g["W"]) < TOL
    if "variance" in g:
        assert relerr(st["variance"], g["variance"]) < TOL


def _ip2_err(W, Wref):
    """IP2 rows are defined up to LAPACK's eigenvector phase."""
    return relerr(phase_align_rows(W, Wref), Wref)


@pytest.mark.parametrize("N", [2, 3, 4])
def test_spatial_kernels_oracle(N):
    g = load("spatial_kernels")
    X, phi, W, U = (g[f"N{N}_{k}"] for k in ("X", "phi", "W", "U"))
    assert relerr(ospatial.weighted_covariance(X, phi), U) < 1e-12
    for fl in ("max", "add", "none"):
        assert relerr(ospatial.update_by_ip1(W, U, FLOORS[fl]), g[f"N{N}_ip1_{fl}"]) < TOL
        assert _ip2_err(ospatial.update_by_ip2(W, U, FLOORS[fl]), g[f"N{N}_ip2_{fl}"]) < TOL
        Y = oilrma.separate(X, W)
        assert relerr(ospatial.update_by_iss1(Y, phi, FLOORS[fl]), g[f"N{N}_iss1_{fl}"]) < TOL
        assert _ip2_err(ospatial.update_by_iss2(Y, phi, FLOORS[fl]), g[f"N{N}_iss2_{fl}"]) < TOL
        assert relerr(ospatial.update_by_ipa(Y, phi, FLOORS[fl]), g[f"N{N}_ipa_{fl}"]) < TOL
    neg = [(m - N, (m + 1) % N - N) for m in range(N)]
    assert _ip2_err(ospatial.update_by_ip2(W, U, pairs=neg), g[f"N{N}_ip2_negpairs"]) < TOL
    import itertools
    comb = list(itertools.combinations(range(N), 2))
    assert _ip2_err(ospatial.update_by_ip2(W, U, pairs=comb), g[f"N{N}_ip2_comb"]) < TOL
    assert _ip2_err(ospatial.update_by_ip2_one_pair(W, U[:, (0, 1)], (0, 1)), g[f"N{N}_ip2pair01"]) < TOL
    assert relerr(ospatial.update_by_ipa(Y, phi, normalization=False, max_iter=5), g[f"N{N}_ipa_nonorm_it5"]) < TOL
    assert relerr(ospatial.update_by_ipa(Y, phi[:, :1, :], max_iter=2), g[f"N{N}_ipa_frameweights"]) < TOL
    # ISS2: rows of the updated pair carry the eigenvector's free phase per bin
    seq = [(m, (m + 1) % N) for m in range(N)]
    assert _ip2_err(ospatial.update_by_iss2(Y, phi, pairs=seq), g[f"N{N}_iss2_seq"]) < TOL
    assert _ip2_err(ospatial.update_by_iss2(Y, phi, pairs=neg), g[f"N{N}_iss2_negpairs"]) < TOL
    assert _ip2_err(ospatial.update_by_iss2(Y, phi, pairs=comb), g[f"N{N}_iss2_comb"]) < TOL


def test_linalg_known_answers():
    """Docstring known answers: ssspy/linalg/inv.py:20-37, ssspy/linalg/eigh.py:53-74,131-152."""
    g = load("linalg")
    np.testing.assert_allclose(olinalg.inv2(g["inv2_in"]),
                               [[[-1.5, 0.5], [1.0, 0.0]], [[-3.5, 2.5], [3.0, -2.0]]], atol=1e-12)
    np.testing.assert_allclose(olinalg.inv2(g["inv2_in"]), g["inv2_out"], atol=1e-12)
    A, B = g["eigh2_A"], g["eigh2_B"]
    lam, z = olinalg.eigh2(A)
    np.testing.assert_allclose(lam, [-0.23606798, 4.23606798], atol=1e-8)
    lam, z = olinalg.eigh2(A, B)
    np.testing.assert_allclose(lam, [-1.61803399, 0.61803399], atol=1e-8)
    np.testing.assert_allclose(A @ z, lam * (B @ z), atol=1e-10)
    for t in (1, 2, 3):
        lam, z = olinalg.eigh2(A, B, type=t)
        np.testing.assert_allclose(lam, g[f"eigh2_lamb_t{t}"], atol=1e-10)
        np.testing.assert_allclose(z, g[f"eigh2_z_t{t}"], atol=1e-10)


@pytest.mark.parametrize("n", [2, 3, 4, 8])
def test_linalg_nxn(n):
    g = load("linalg")
    a, rhs, Ah, Bh = g[f"n{n}_a"], g[f"n{n}_rhs"], g[f"n{n}_A"], g[f"n{n}_B"]
    assert relerr(olinalg.solve(a, rhs), g[f"n{n}_solve"]) < 1e-12
    for t in (1, 2, 3):
        lam, z = olinalg.eigh(Ah, Bh, type=t)
        np.testing.assert_allclose(lam, g[f"n{n}_lamb_t{t}"], rtol=1e-10)
        lhs = {1: Ah @ z, 2: Ah @ Bh @ z, 3: Bh @ Ah @ z}[t]
        rhs_ = {1: (Bh @ z) * lam[:, None, :], 2: z * lam[:, None, :], 3: z * lam[:, None, :]}[t]
        assert relerr(lhs, rhs_) < 1e-9
    for ref, key in ((0, "ref0"), (1, "ref1"), (None, "refnone")):
        assert relerr(projection_back(a, reference_id=ref), g[f"n{n}_pb_w_{key}"]) < 1e-12


def test_projection_back_y_form():
    g = load("linalg")
    for ref, key in ((0, "ref0"), (2, "ref2"), (None, "refnone")):
        out = projection_back(g["pb_Y"], reference=g["pb_X"], reference_id=ref)
        assert out.shape == g[f"pb_y_{key}"].shape
        assert relerr(out, g[f"pb_y_{key}"]) < 1e-12


@pytest.mark.parametrize("name", golden_cases("mnmf_"))
def test_fast_gauss_mnmf_oracle_matches_reference(name):
    from oracle import mnmf as omnmf
    g = load(name)
    st = omnmf.run(g["X"], g["T0"], g["V0"], g["Q0"], g["D0"], int(g["n_iter"]), floor=FLOORS[str(g["flooring"])],
                   algorithm=str(g["algorithm"]), pairs=[tuple(p) for p in g["pairs"]],
                   normalization=bool(g["normalization"]), reference_id=int(g["reference_id"]))
    np.testing.assert_allclose(st["loss"], g["loss"], rtol=1e-10, atol=1e-9)
    assert relerr(st["T"], g["T"]) < TOL and relerr(st["V"], g["V"]) < TOL and relerr(st["D"], g["D"]) < TOL
    Q = st["Q"]
    if str(g["algorithm"]) == "IP2":
        Q = phase_align_rows(Q, g["Q"])
    assert relerr(Q, g["Q"]) < TOL
    assert relerr(st["Y"], g["Y"]) < 1e-8


def test_minimal_distortion_principle_oracle():
    from oracle.projection_back import minimal_distortion_principle
    g = load("mdp")
    for ref, key in ((0, "ref0"), (2, "ref2"), (None, "refnone")):
        assert relerr(minimal_distortion_principle(g["Y"], g["X"], ref), g["mdp_" + key]) < 1e-12


@pytest.mark.parametrize("name", golden_cases("fdica_"))
def test_fdica_oracle_matches_reference(name):
    """AuxLaplaceFDICA (ssspy/bss/fdica.py:846-1245, :1527-1667) incl. the permutation alignment and both scale
    restorations, against fixtures generated by the unmodified reference (tests/golden/make_golden_fdica.py)."""
    from oracle import fdica as ofdica
    g = load(name)
    spatial = str(g["spatial"])
    st = ofdica.run(g["X"], int(g["n_iter"]), W=g.get("W0"), floor=FLOORS[str(g["flooring"])], spatial_algorithm=spatial,
                    pairs=[tuple(p) for p in g["pairs"]] if spatial == "IP2" else None,
                    reference_id=int(g["reference_id"]), permutation_alignment=bool(g["permutation_alignment"]),
                    scale_restoration=sr_arg(g["scale_restoration"]))
    np.testing.assert_allclose(st["loss"], g["loss"], rtol=1e-9)
    # IP2 on few frames: the 2x2 generalised eigenvectors are ill-conditioned (DESIGN.md "IP2 sensitivity"); the
    # closed-form solver of the oracle and LAPACK then agree to 2e-7 on Y while the loss agrees to 1e-9
    tol = 1e-6 if spatial == "IP2" else TOL
    if sr_arg(g["scale_restoration"]) or spatial != "IP2":
        assert relerr(st["Y"], g["Y"]) < tol
        assert _ip2_err(st["W"], g["W"]) < max(tol, 1e-8)
    else:
        assert relerr(np.abs(st["Y"]), np.abs(g["Y"])) < TOL


@pytest.mark.parametrize("N", [2, 3, 4])
def test_permutation_solver_oracle(N):
    """correlation_based_permutation_solver (ssspy/algorithm/permutation_alignment.py:12-121): outputs are
    permutations of the inputs, so the comparison is exact."""
    from oracle import fdica as ofdica
    g = load("permutation_solver")
    Y, W = g[f"N{N}_Y"], g[f"N{N}_W"]
    Yo, Wo, order, perms = ofdica.correlation_based_permutation_solver(Y, W)
    np.testing.assert_array_equal(Yo, g[f"N{N}_Yout"])
    np.testing.assert_array_equal(Wo, g[f"N{N}_Wout"])
    assert sorted(order.tolist()) == list(range(Y.shape[0])) and perms.shape == (Y.shape[0], N)
    Ya = ofdica.correlation_based_permutation_solver(Y, floor=FLOORS["add"])[0]
    np.testing.assert_array_equal(Ya, g[f"N{N}_Yout_add"])


@pytest.mark.parametrize("normalization", [True, False])
def test_deferred_power_normalisation_commutes_with_the_source_model(normalization):
    """The identity behind the fused covariance + IP1 + basis kernel (ssb_fused_spatial_source, SSB_TMA bit 2): n x update_once (ilrma.py:900-922) equals
    [T, V]_1, { [U, IP1]_t, [T, V]_(t+1) with the UNNORMALISED W_t and T_t, then T /= psi_t^2, W /= psi_t }, [U, IP1,
    normalise]_n -- the ratio of the basis update and the whole activation update are invariant under
    (P, T) -> (P, T) / psi^2.  Checked in fp64 with the oracle."""
    from oracle import ilrma as o
    from ssspy_b200.utils.synth import make_batch, make_nmf_init
    N, I, J, K, n_iter = 2, 37, 48, 5, 5
    X = make_batch(1, N, I, J, config_id=23, mode="mix")[0]
    T, V = make_nmf_init(N, I, J, K, seed=11)
    ref = o.run(X, T, V, n_iter, normalization=normalization, record_loss=False, scale_restoration=False)
    st = o.init_state(X, T, V, None, "IP", None)
    o.update_basis(st)
    o.update_activation(st)
    for _ in range(1, n_iter):
        o.update_spatial(st)
        psi = np.maximum(np.sqrt(np.mean(np.abs(o.separate(st["X"], st["W"])) ** 2, axis=(-2, -1))), 1e-10)
        o.update_basis(st)
        o.update_activation(st)
        if normalization:
            st["T"] = st["T"] / psi[:, None, None] ** 2
            st["W"] = st["W"] / psi[None, :, None]
    o.update_spatial(st)
    if normalization:
        o.normalize(st)
    for k in ("T", "V", "W"):
        assert np.abs(st[k] - ref[k]).max() <= 1e-12 * np.abs(ref[k]).max()


def test_linalg_operators_of_the_ipa_path_match_reference():
    """cbrt, solve_cubic and lqpqm2 (ssspy/linalg/cubic.py:4, polynomial.py:9, lqpqm.py:13) against vectors produced
    by the unmodified reference (tests/golden/make_golden_linalg_ops.py)."""
    g = load("linalg_ops")
    np.testing.assert_allclose(olinalg.cbrt(g["cbrt_real_in"]), g["cbrt_real_out"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(olinalg.cbrt(g["cbrt_cplx_in"]), g["cbrt_cplx_out"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(olinalg.solve_cubic(g["cubic_A"], g["cubic_B"], g["cubic_C"]), g["cubic_roots"],
                               rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(olinalg.solve_cubic(g["cubic_A"], g["cubic_B"], g["cubic_C"], all=False),
                               g["cubic_first"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(olinalg.solve_cubic(g["cubic_cA"], g["cubic_cB"], g["cubic_cC"]), g["cubic_croots"],
                               rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(olinalg.solve_cubic(g["cubic_gA"], g["cubic_gB"], g["cubic_gC"], g["cubic_gD"]),
                               g["cubic_groots"], rtol=1e-12, atol=1e-13)
    with pytest.raises(np.linalg.LinAlgError, match="Coefficients include zero"):
        olinalg.solve_cubic(np.array([1.0, 0.0]), np.ones(2), np.ones(2), np.ones(2))
    for M in (1, 2, 3, 5):
        for it in (1, 10):
            y = olinalg.lqpqm2(g["lqpqm2_M%d_H" % M], g["lqpqm2_M%d_v" % M], g["lqpqm2_M%d_z" % M], FLOORS["max"],
                               max_iter=it)
            assert relerr(y, g["lqpqm2_M%d_it%d_out" % (M, it)]) < 1e-9, (M, it)


def test_stft_oracle_matches_scipy():
    """oracle.transform against scipy.signal.stft / istft called as the reference's notebooks call them
    (tests/golden/make_golden_stft.py); incl. an inverse of a spectrogram that is not the STFT of any signal."""
    from oracle import transform as ot
    g = load("stft")
    for c in "abcd":
        n, h = int(g[c + "_n_fft"]), int(g[c + "_hop"])
        w = ot.hann(n)
        np.testing.assert_allclose(ot.stft(g[c + "_x"], w, h), g[c + "_Z"], atol=1e-13)
        np.testing.assert_allclose(ot.istft(g[c + "_Z"], w, h), g[c + "_y"], atol=1e-12)
        np.testing.assert_allclose(ot.istft(g[c + "_Zr"], w, h), g[c + "_yr"], atol=1e-12)
    np.testing.assert_allclose(ot.stft(g["w_x"], g["w_win"], 32), g["w_Z"], atol=1e-13)
    np.testing.assert_allclose(ot.istft(g["w_Z"], g["w_win"], 32), g["w_y"], atol=1e-12)
