"""Generate golden vectors for the hot path by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The GPU box has no /root/reference; tests read only
the committed .npz files.  Inputs are stored in the fixtures (complex128) so the
consumer never needs the reference.
"""
import functools
import os
import sys

import numpy as np

REF = os.environ.get("SSSPY_REF", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from ssspy.algorithm import projection_back  # noqa: E402
from ssspy.bss._update_spatial_model import (  # noqa: E402
    update_by_ip1, update_by_ip2, update_by_ip2_one_pair, update_by_ipa, update_by_iss1, update_by_iss2)
from ssspy.bss.ilrma import GGDILRMA, TILRMA, GaussILRMA  # noqa: E402
from ssspy.bss.iva import AuxGaussIVA, AuxLaplaceIVA  # noqa: E402
from ssspy.bss.mnmf import FastGaussMNMF  # noqa: E402
from ssspy.linalg import eigh, eigh2, inv2  # noqa: E402
from ssspy.linalg._solve import solve  # noqa: E402
from ssspy.special.flooring import add_flooring, max_flooring  # noqa: E402
from ssspy.utils.select_pair import combination_pair_selector, sequential_pair_selector  # noqa: E402

from ssspy_b200.utils.synth import make_mixture, make_nmf_init  # noqa: E402

FLOOR = {"max": functools.partial(max_flooring, eps=1e-10),
         "add": functools.partial(add_flooring, eps=1e-10), "none": None}


def rand_w(rng, I, N):
    """Non-identity, well-conditioned initial demixing filters."""
    return np.eye(N)[None] + 0.3 * (rng.standard_normal((I, N, N)) + 1j * rng.standard_normal((I, N, N)))


def ilrma_case(name, N, I, J, K, n_iter, spatial="IP", source="MM", domain=2, normalization=True,
               flooring="max", reference_id=0, scale_restoration=True, pairs=None, w_init=False, seed=0,
               dist="gauss", dist_param=0.0, partitioning=False, ipa=None):
    X = make_mixture(N, I, J, seed=seed, mode="mix")
    T, V = make_nmf_init(N, I, J, K, seed=42 + seed)
    kwargs = dict(basis=T, activation=V)
    Z0 = None
    if partitioning:  # shared T[I,K], V[K,J] + latent Z[N,K] with unit column sums (ilrma.py:219-245)
        prng = np.random.default_rng(99 + seed)
        Z0 = prng.random((N, K)) + 0.1
        Z0 = Z0 / Z0.sum(axis=0)
        T, V = T[0].copy(), V[0].copy()
        kwargs = dict(basis=T, activation=V, latent=Z0)
    rng = np.random.default_rng(7 + seed)
    W0 = rand_w(rng, I, N) if w_init else None
    if W0 is not None:
        kwargs["demix_filter"] = W0
    snaps = []

    def cb(m):
        snaps.append((None if m.demix_filter is None else m.demix_filter.copy(), m.output.copy(),
                      m.basis.copy(), m.activation.copy()))
    sel = None
    if pairs == "combination":
        sel = combination_pair_selector
    elif pairs == "sequential_sorted":
        sel = functools.partial(sequential_pair_selector, sort=True)
    common = dict(spatial_algorithm=spatial, source_algorithm=source, domain=domain, flooring_fn=FLOOR[flooring],
                  partitioning=partitioning,
                  pair_selector=sel, callbacks=cb, normalization=normalization, scale_restoration=scale_restoration,
                  record_loss=True, reference_id=reference_id, rng=np.random.default_rng(0))
    if ipa is not None:
        common.update(lqpqm_normalization=ipa[0], newton_iter=ipa[1])
    if dist == "t":
        m = TILRMA(n_basis=K, dof=dist_param, **common)
    elif dist == "ggd":
        m = GGDILRMA(n_basis=K, beta=dist_param, **common)
    else:
        m = GaussILRMA(n_basis=K, **common)
    Y = m(X, n_iter=n_iter, **kwargs)
    pair_list = np.array(list((sel or sequential_pair_selector)(N)), dtype=np.int32)
    out = dict(kind="ilrma", X=X, T0=T, V0=V, Y=Y, T=m.basis, V=m.activation, loss=np.array(m.loss),
               Y_last_iter=snaps[-1][1], T_first_iter=snaps[1][2], V_first_iter=snaps[1][3],
               n_iter=n_iter, spatial=spatial, source=source, domain=float(domain),
               normalization=str(normalization), flooring=flooring,
               reference_id=-1 if reference_id is None else reference_id,
               scale_restoration=scale_restoration, pairs=pair_list, dist=dist, dist_param=float(dist_param),
               ipa_normalization=True if ipa is None else bool(ipa[0]), ipa_newton_iter=1 if ipa is None else int(ipa[1]))
    if W0 is not None:
        out["W0"] = W0
    if partitioning:
        out["Z0"], out["Z"] = Z0, m.latent
    if m.demix_filter is not None:
        out["W"] = m.demix_filter
        out["W_first_iter"] = snaps[1][0]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", m.loss[0], "->", m.loss[-1])


def iva_case(name, N, I, J, n_iter, model="laplace", spatial="IP", flooring="max", reference_id=0,
             scale_restoration=True, pairs=None, w_init=False, seed=0, ipa=None):
    X = make_mixture(N, I, J, seed=100 + seed, mode="mix")
    rng = np.random.default_rng(9 + seed)
    kwargs = {}
    W0 = rand_w(rng, I, N) if w_init else None
    if W0 is not None:
        kwargs["demix_filter"] = W0
    sel = combination_pair_selector if pairs == "combination" else None
    cls = AuxLaplaceIVA if model == "laplace" else AuxGaussIVA
    extra = {} if ipa is None else dict(lqpqm_normalization=ipa[0], newton_iter=ipa[1])
    m = cls(spatial_algorithm=spatial, flooring_fn=FLOOR[flooring], pair_selector=sel,
            scale_restoration=scale_restoration, record_loss=True, reference_id=reference_id, **extra)
    Y = m(X, n_iter=n_iter, **kwargs)
    pair_list = np.array(list((sel or sequential_pair_selector)(N)), dtype=np.int32)
    out = dict(kind="iva", X=X, Y=Y, loss=np.array(m.loss), n_iter=n_iter, model=model, spatial=spatial,
               flooring=flooring, reference_id=reference_id, scale_restoration=scale_restoration,
               pairs=pair_list, ipa_normalization=True if ipa is None else bool(ipa[0]),
               ipa_newton_iter=1 if ipa is None else int(ipa[1]))
    if W0 is not None:
        out["W0"] = W0
    if m.demix_filter is not None:
        out["W"] = m.demix_filter
    if model == "gauss":
        out["variance"] = m.variance
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", m.loss[0], "->", m.loss[-1])


def mnmf_case(name, N, I, J, K, n_iter, algorithm="IP", flooring="max", reference_id=0, pairs=None, normalization=True,
              seed=0):
    """FastGaussMNMF with injected state (tests/regression/bss/test_mnmf.py:108-133 pattern)."""
    X = make_mixture(N, I, J, seed=200 + seed, mode="mix")
    rng = np.random.default_rng(300 + seed)
    T = rng.random((N, I, K))
    V = rng.random((N, K, J))
    D = rng.random((I, N, N))
    Q = rand_w(rng, I, N) if seed % 2 else np.tile(np.eye(N, dtype=np.complex128), (I, 1, 1))
    sel = combination_pair_selector if pairs == "combination" else None
    m = FastGaussMNMF(n_basis=K, n_sources=N, diagonalizer_algorithm=algorithm, flooring_fn=FLOOR[flooring],
                      pair_selector=sel, normalization=normalization, record_loss=True, reference_id=reference_id,
                      rng=np.random.default_rng(0))
    Y = m(X, n_iter=n_iter, basis=T, activation=V, spatial=D, diagonalizer=Q)
    pair_list = np.array(list((sel or sequential_pair_selector)(N)), dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), kind="mnmf", X=X, T0=T, V0=V, D0=D, Q0=Q, Y=Y,
                        T=m.basis, V=m.activation, D=m.spatial, Q=m.diagonalizer, loss=np.array(m.loss),
                        n_iter=n_iter, algorithm=algorithm, flooring=flooring, reference_id=reference_id,
                        normalization=normalization, pairs=pair_list)
    print(name, "loss", m.loss[0], "->", m.loss[-1])


def kernel_cases():
    """update_by_* on the reference's own smoke-test shapes (tests/package/bss/
    test_update_spatial_model.py:45-171: rng=default_rng(42), (I,J)=(31,20), N in {2,3})."""
    out = {}
    for N in (2, 3, 4):
        rng = np.random.default_rng(42)
        I, J = 31, 20
        X = rng.standard_normal((N, I, J)) + 1j * rng.standard_normal((N, I, J))
        phi = 1 / (rng.random((N, I, J)) + 0.1)
        W = rand_w(rng, I, N)
        XX = X[:, None] * X[None].conj()
        U = np.mean(phi.transpose(1, 0, 2)[:, :, None, None, :] * XX.transpose(2, 0, 1, 3)[:, None], axis=-1)
        out[f"N{N}_X"], out[f"N{N}_phi"], out[f"N{N}_W"], out[f"N{N}_U"] = X, phi, W, U
        for fl in ("max", "add", "none"):
            out[f"N{N}_ip1_{fl}"] = update_by_ip1(W, U, flooring_fn=FLOOR[fl], overwrite=False)
            out[f"N{N}_ip2_{fl}"] = update_by_ip2(W, U, flooring_fn=FLOOR[fl], overwrite=False)
            Y = (W @ X.transpose(1, 0, 2)).transpose(1, 0, 2)
            out[f"N{N}_iss1_{fl}"] = update_by_iss1(Y, phi, flooring_fn=FLOOR[fl])
            out[f"N{N}_iss2_{fl}"] = update_by_iss2(Y, phi, flooring_fn=FLOOR[fl])
            out[f"N{N}_ipa_{fl}"] = update_by_ipa(Y, phi, flooring_fn=FLOOR[fl])
        # pair selector with negative indices (test_update_spatial_model.py:19-24)
        def neg_sel(n):
            for m in range(n):
                yield m - n, (m + 1) % n - n
        out[f"N{N}_ip2_negpairs"] = update_by_ip2(W, U, pair_selector=neg_sel, overwrite=False)
        out[f"N{N}_ip2_comb"] = update_by_ip2(W, U, pair_selector=combination_pair_selector, overwrite=False)
        out[f"N{N}_ip2pair01"] = update_by_ip2_one_pair(W, U[:, (0, 1)], pair=(0, 1))
        out[f"N{N}_ipa_nonorm_it5"] = update_by_ipa(Y, phi, normalization=False, max_iter=5)
        out[f"N{N}_ipa_frameweights"] = update_by_ipa(Y, phi[:, :1, :].transpose(0, 1, 2), max_iter=2)
        out[f"N{N}_iss2_seq"] = update_by_iss2(Y, phi, pair_selector=sequential_pair_selector)  # incl. (N-1, 0)
        out[f"N{N}_iss2_negpairs"] = update_by_iss2(Y, phi, pair_selector=neg_sel)
        out[f"N{N}_iss2_comb"] = update_by_iss2(Y, phi, pair_selector=combination_pair_selector)
    np.savez_compressed(os.path.join(HERE, "spatial_kernels.npz"), **out)
    print("spatial_kernels done")


def linalg_cases():
    rng = np.random.default_rng(111)
    out = {}
    out["inv2_in"] = np.array([[[0, 1], [2, 3]], [[4, 5], [6, 7]]], dtype=np.float64)
    out["inv2_out"] = inv2(out["inv2_in"])
    A = np.array([[1, -2j], [2j, 3]])
    B = np.array([[2, -3j], [3j, 5]])
    out["eigh2_A"], out["eigh2_B"] = A, B
    out["eigh2_lamb_std"], _ = eigh2(A)
    for t in (1, 2, 3):
        lam, z = eigh2(A, B, type=t)
        out[f"eigh2_lamb_t{t}"], out[f"eigh2_z_t{t}"] = lam, z
    for n in (2, 3, 4, 8):
        a = rng.standard_normal((16, n, n)) + 1j * rng.standard_normal((16, n, n))
        b = rng.standard_normal((16, n, n)) + 1j * rng.standard_normal((16, n, n))
        Ah = a @ a.conj().transpose(0, 2, 1)
        Bh = b @ b.conj().transpose(0, 2, 1) + n * np.eye(n)
        rhs = rng.standard_normal((16, n)) + 1j * rng.standard_normal((16, n))
        out[f"n{n}_A"], out[f"n{n}_B"], out[f"n{n}_a"], out[f"n{n}_rhs"] = Ah, Bh, a, rhs
        out[f"n{n}_solve"] = solve(a, rhs)
        out[f"n{n}_inv"] = np.linalg.inv(a)
        for t in (1, 2, 3):
            lam, z = eigh(Ah, Bh, type=t)
            out[f"n{n}_lamb_t{t}"] = lam
        out[f"n{n}_lamb_std"], _ = eigh(Ah)
        W = a
        out[f"n{n}_pb_w_ref0"] = projection_back(W, reference_id=0)
        out[f"n{n}_pb_w_ref1"] = projection_back(W, reference_id=1)
        out[f"n{n}_pb_w_refnone"] = projection_back(W, reference_id=None)
    X = make_mixture(3, 9, 14, seed=5)
    Y = make_mixture(3, 9, 14, seed=6)
    out["pb_X"], out["pb_Y"] = X, Y
    out["pb_y_ref0"] = projection_back(Y, reference=X, reference_id=0)
    out["pb_y_ref2"] = projection_back(Y, reference=X, reference_id=2)
    out["pb_y_refnone"] = projection_back(Y, reference=X, reference_id=None)
    np.savez_compressed(os.path.join(HERE, "linalg.npz"), **out)
    print("linalg done")


def tggd_cases():
    """TILRMA (ssspy/bss/ilrma.py:1992-3334) and GGDILRMA (:3337-4410)."""
    ilrma_case("tilrma_ip1_mm", 3, 17, 23, 4, 5, dist="t", dist_param=5.0, seed=30)
    ilrma_case("tilrma_ip1_mm_p1_n2", 2, 21, 30, 3, 6, domain=1, dist="t", dist_param=100.0, seed=31)
    ilrma_case("tilrma_ip1_me", 3, 17, 23, 4, 5, source="ME", dist="t", dist_param=10.0, seed=32)
    ilrma_case("tilrma_ip2_mm", 3, 17, 23, 4, 5, spatial="IP2", dist="t", dist_param=5.0, w_init=True, seed=33)
    ilrma_case("tilrma_iss1_mm", 3, 17, 23, 4, 5, spatial="ISS", dist="t", dist_param=5.0, seed=34)
    ilrma_case("ggdilrma_ip1_b1", 3, 17, 23, 4, 5, dist="ggd", dist_param=1.0, seed=35)
    ilrma_case("ggdilrma_ip1_b19_p1", 2, 21, 30, 3, 6, domain=1, dist="ggd", dist_param=1.9, seed=36)
    ilrma_case("ggdilrma_ip2_b15", 3, 20, 48, 5, 4, spatial="IP2", dist="ggd", dist_param=1.5, seed=37)
    ilrma_case("ggdilrma_iss1_b1_pbnorm", 3, 17, 23, 4, 5, spatial="ISS", normalization="projection_back", dist="ggd",
               dist_param=1.0, seed=38)


def partitioning_cases():
    """partitioning=True (latent Z; ilrma.py:201-245, :1007-1049, :1098-1113, :1174-1189, :424-430)."""
    ilrma_case("ilrma_part_ip1_mm", 3, 17, 23, 4, 5, partitioning=True, seed=40)
    ilrma_case("ilrma_part_ip2_me_n2", 2, 21, 30, 5, 5, spatial="IP2", source="ME", partitioning=True, seed=41)
    ilrma_case("ilrma_part_iss1_p1", 3, 17, 23, 4, 5, spatial="ISS", domain=1, partitioning=True, seed=42)
    ilrma_case("ilrma_part_ip1_nonorm", 4, 12, 40, 6, 4, normalization=False, partitioning=True, seed=43)
    ilrma_case("tilrma_part_ip1_mm", 3, 17, 23, 4, 5, dist="t", dist_param=4.0, partitioning=True, seed=44)
    ilrma_case("tilrma_part_iss1_me", 3, 17, 23, 4, 5, spatial="ISS", source="ME", dist="t", dist_param=20.0,
               partitioning=True, seed=45)
    ilrma_case("ggdilrma_part_ip1_b1", 3, 17, 23, 4, 5, dist="ggd", dist_param=1.0, partitioning=True, seed=46)


def iss2_cases():
    """spatial_algorithm="ISS2" through the classes (ilrma.py:1698-1811, iva.py:1968-2066): default pair selector =
    sequential (all N pairs incl. the wrapping (N-1, 0))."""
    ilrma_case("ilrma_iss2", 3, 17, 23, 4, 5, spatial="ISS2", seed=50)
    ilrma_case("ilrma_iss2_n2_p1", 2, 21, 30, 3, 5, spatial="ISS2", domain=1, seed=51)
    ilrma_case("ilrma_iss2_n4_comb_pb", 4, 12, 40, 5, 4, spatial="ISS2", pairs="combination",
               normalization="projection_back", seed=52)
    ilrma_case("tilrma_iss2", 3, 17, 23, 4, 5, spatial="ISS2", dist="t", dist_param=6.0, seed=53)
    ilrma_case("ggdilrma_iss2_part", 3, 17, 23, 4, 5, spatial="ISS2", dist="ggd", dist_param=1.2, partitioning=True, seed=54)
    iva_case("iva_laplace_iss2", 3, 17, 23, 6, spatial="ISS2", seed=55)
    iva_case("iva_gauss_iss2_n4", 4, 12, 40, 5, model="gauss", spatial="ISS2", seed=56)
    iva_case("iva_laplace_iss2_n2_comb", 2, 21, 30, 6, spatial="ISS2", pairs="combination", seed=57)


def ipa_cases():
    """spatial_algorithm="IPA" (ilrma.py:1813-1908, iva.py:2068-2176)."""
    ilrma_case("ilrma_ipa", 3, 17, 23, 4, 5, spatial="IPA", seed=60)
    ilrma_case("ilrma_ipa_n2_p1_it3", 2, 21, 30, 3, 5, spatial="IPA", domain=1, ipa=(True, 3), seed=61)
    ilrma_case("ilrma_ipa_n4_nonorm_part", 4, 12, 40, 5, 4, spatial="IPA", ipa=(False, 2), partitioning=True, seed=62)
    iva_case("iva_laplace_ipa", 3, 17, 23, 6, spatial="IPA", seed=63)
    iva_case("iva_gauss_ipa_n4", 4, 12, 40, 5, model="gauss", spatial="IPA", ipa=(True, 2), seed=64)
    iva_case("iva_laplace_ipa_n2", 2, 21, 30, 6, spatial="IPA", ipa=(False, 1), seed=65)


def mdp_cases():
    """minimal_distortion_principle standalone + as scale_restoration of GaussILRMA / AuxLaplaceIVA."""
    from ssspy.algorithm import minimal_distortion_principle
    out = {}
    X = make_mixture(3, 9, 14, seed=5)
    Y = make_mixture(3, 9, 14, seed=6)
    out["X"], out["Y"] = X, Y
    out["mdp_ref0"] = minimal_distortion_principle(Y, reference=X, reference_id=0)
    out["mdp_ref2"] = minimal_distortion_principle(Y, reference=X, reference_id=2)
    out["mdp_refnone"] = minimal_distortion_principle(Y, reference=X, reference_id=None)
    np.savez_compressed(os.path.join(HERE, "mdp.npz"), **out)
    ilrma_case("ilrma_ip1_mdp", 3, 17, 23, 4, 5, scale_restoration="minimal_distortion_principle", reference_id=1, seed=20)
    ilrma_case("ilrma_iss1_mdp", 3, 17, 23, 4, 5, spatial="ISS", scale_restoration="MDP", seed=21)
    print("mdp done")


def main():
    linalg_cases()
    kernel_cases()
    mdp_cases()
    tggd_cases()
    partitioning_cases()
    iss2_cases()
    ipa_cases()
    # GaussILRMA: spatial x source x domain x normalisation x flooring grid (regression-test pattern,
    # tests/regression/bss/test_ilrma.py:48-62: inject basis/activation, fixed n_iter, compare).
    ilrma_case("ilrma_ip1_mm_n2", 2, 33, 40, 4, 10)
    ilrma_case("ilrma_ip1_mm_n3_winit", 3, 17, 23, 4, 5, w_init=True, seed=1)
    ilrma_case("ilrma_ip1_mm_n4", 4, 20, 36, 5, 8, seed=2)
    ilrma_case("ilrma_ip1_mm_n8", 8, 9, 160, 3, 4, seed=3)
    ilrma_case("ilrma_ip1_mm_p1", 3, 17, 23, 4, 5, domain=1, seed=4)
    ilrma_case("ilrma_ip1_me", 3, 17, 23, 4, 5, source="ME", seed=5)
    ilrma_case("ilrma_ip1_nonorm", 2, 17, 23, 4, 5, normalization=False, seed=6)
    ilrma_case("ilrma_ip1_pbnorm", 3, 17, 23, 4, 5, normalization="projection_back", reference_id=1, seed=7)
    ilrma_case("ilrma_ip1_addfloor", 2, 17, 23, 4, 5, flooring="add", seed=8)
    ilrma_case("ilrma_ip1_nofloor_noscale", 2, 17, 23, 4, 5, flooring="none", scale_restoration=False, seed=9)
    ilrma_case("ilrma_ip2_mm_n2", 2, 33, 40, 4, 10, spatial="IP2", seed=10)
    ilrma_case("ilrma_ip2_mm_n3", 3, 17, 23, 4, 5, spatial="IP2", w_init=True, seed=11)
    ilrma_case("ilrma_ip2_mm_n4_comb", 4, 20, 36, 5, 6, spatial="IP2", pairs="combination", seed=12)
    ilrma_case("ilrma_ip2_p1", 3, 17, 23, 4, 5, spatial="IP2", domain=1, seed=13)
    ilrma_case("ilrma_ip2_n8", 8, 9, 160, 3, 3, spatial="IP2", seed=14)
    ilrma_case("ilrma_iss1_mm_n2", 2, 33, 40, 4, 10, spatial="ISS", seed=15)
    ilrma_case("ilrma_iss1_mm_n3", 3, 17, 23, 4, 5, spatial="ISS", reference_id=2, seed=16)
    ilrma_case("ilrma_iss1_mm_n4_p1", 4, 20, 36, 5, 6, spatial="ISS", domain=1, seed=17)
    ilrma_case("ilrma_iss1_pbnorm", 3, 17, 23, 4, 5, spatial="ISS", normalization="projection_back", seed=18)
    ilrma_case("ilrma_iss1_me_nonorm", 3, 17, 23, 4, 5, spatial="ISS", source="ME", normalization=False, seed=19)
    # AuxIVA
    for model in ("laplace", "gauss"):
        iva_case(f"iva_{model}_ip1_n2", 2, 33, 40, 10, model=model)
        iva_case(f"iva_{model}_ip1_n3_winit", 3, 17, 23, 5, model=model, w_init=True, seed=1)
        iva_case(f"iva_{model}_ip2_n2", 2, 33, 40, 8, model=model, spatial="IP2", seed=2)
        iva_case(f"iva_{model}_ip2_n4_comb", 4, 20, 36, 5, model=model, spatial="IP2", pairs="combination", seed=3)
        iva_case(f"iva_{model}_iss1_n2", 2, 33, 40, 10, model=model, spatial="ISS", seed=4)
        iva_case(f"iva_{model}_iss1_n4", 4, 20, 36, 6, model=model, spatial="ISS", reference_id=1, seed=5)
    iva_case("iva_laplace_ip1_addfloor_noscale", 3, 17, 23, 5, flooring="add", scale_restoration=False, seed=6)
    iva_case("iva_laplace_ip1_n8", 8, 9, 160, 4, seed=7)
    # FastGaussMNMF (config 5 family)
    mnmf_case("mnmf_ip1_n2", 2, 17, 24, 3, 5)
    mnmf_case("mnmf_ip1_n3_qinit", 3, 13, 20, 4, 4, seed=1)
    mnmf_case("mnmf_ip1_n4_ref1", 4, 11, 32, 3, 4, reference_id=1, seed=2)
    mnmf_case("mnmf_ip2_n2", 2, 17, 24, 3, 5, algorithm="IP2", seed=3)
    mnmf_case("mnmf_ip2_n4_comb", 4, 11, 32, 3, 4, algorithm="IP2", pairs="combination", seed=4)
    mnmf_case("mnmf_ip1_addfloor_nonorm", 3, 13, 20, 4, 4, flooring="add", normalization=False, seed=5)


if __name__ == "__main__":
    main()
