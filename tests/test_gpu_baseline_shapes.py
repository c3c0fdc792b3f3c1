"""GPU parity at the BASELINE.json shapes: the CUDA path through the separator classes and the
default chunked engine (>= 2 chunk plans on their own streams) against the fp64 oracle, >= 5 iterations, every
mixture of the batch.  Same tolerances as tests/test_gpu_parity.py: Y (after projection back) rel-Frobenius <= 1e-4,
T / V <= 1e-4, loss trajectory rel <= 1e-5.  The oracle runs in seconds to a couple of minutes per case (config 4 is
the largest: N = 8, I = 2049, J = 1024, K = 32)."""
import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _loss_close(got, want, rtol=1e-5):
    np.testing.assert_allclose(np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64), rtol=rtol,
                               atol=1e-4)


@pytest.mark.parametrize("N,spatial,I,J,K,n_iter", [
    (2, "IP", 1025, 512, 16, 5),     # configs[1]
    (4, "IP", 1025, 512, 16, 5),     # north_star: n_channels in {2, 4, 8} at n_bins = 1025, n_frames = 512
    (8, "IP", 1025, 512, 16, 5),
    (2, "IP2", 1025, 512, 16, 5),
    (4, "IP2", 1025, 512, 16, 5),
    (4, "ISS", 1025, 512, 16, 5),
    (8, "IP2", 2049, 1024, 32, 5),   # configs[3] (one GPU's shard is 64 of these mixtures)
])
def test_gauss_ilrma_at_baseline_shapes(N, spatial, I, J, K, n_iter):
    from oracle import ilrma as oilrma
    from ssspy_b200.bss import GaussILRMA
    from ssspy_b200.utils.synth import make_batch, make_nmf_init
    B = 2
    X = make_batch(B, N, I, J, config_id=40 + N, mode="mix")
    T, V = make_nmf_init(N, I, J, K, seed=42)
    m = GaussILRMA(n_basis=K, spatial_algorithm=spatial)
    m.chunk_size = 1  # two chunk plans on two streams: the engine path bench.py times
    Y = m(X, n_iter=n_iter, basis=T, activation=V)
    assert len(m._chunks) == 2
    loss = np.asarray(m.loss)
    assert loss.shape == (n_iter + 1, B) and np.all(np.isfinite(loss))
    for b in range(B):
        st = oilrma.run(X[b], T, V, n_iter, spatial_algorithm=spatial)
        ey, et, ev = relerr(Y[b], st["Y"]), relerr(m.basis[b], st["T"]), relerr(m.activation[b], st["V"])
        print("GaussILRMA-%s N=%d I=%d J=%d K=%d mixture %d: relerr Y %.2e T %.2e V %.2e loss %.2e" % (
            spatial, N, I, J, K, b, ey, et, ev, np.max(np.abs(loss[:, b] / np.asarray(st["loss"]) - 1))))
        assert ey < TOL and et < TOL and ev < TOL
        _loss_close(loss[:, b], st["loss"])


@pytest.mark.parametrize("cls_name,spatial,N,n_iter", [("AuxLaplaceIVA", "ISS", 4, 5),   # configs[2]
                                                       ("AuxLaplaceIVA", "IP", 4, 5), ("AuxGaussIVA", "IP2", 4, 5),
                                                       ("AuxLaplaceIVA", "IP", 2, 5)])
def test_aux_iva_at_baseline_shapes(cls_name, spatial, N, n_iter):
    from oracle import iva as oiva
    from ssspy_b200 import bss
    from ssspy_b200.utils.synth import make_batch
    B, I, J = 2, 1025, 512
    X = make_batch(B, N, I, J, config_id=30 + N, mode="mix")
    m = getattr(bss, cls_name)(spatial_algorithm=spatial)
    m.chunk_size = 1
    Y = m(X, n_iter=n_iter)
    assert len(m._chunks) == 2
    model = "laplace" if cls_name == "AuxLaplaceIVA" else "gauss"
    loss = np.asarray(m.loss)
    for b in range(B):
        st = oiva.run(X[b], n_iter, spatial_algorithm=spatial, model=model)
        ey = relerr(Y[b], st["Y"])
        print("%s-%s N=%d mixture %d: relerr Y %.2e loss %.2e" % (
            cls_name, spatial, N, b, ey, np.max(np.abs(loss[:, b] / np.asarray(st["loss"]) - 1))))
        assert ey < TOL
        _loss_close(loss[:, b], st["loss"])


@pytest.mark.parametrize("alg", ["IP", "IP2"])
def test_fast_gauss_mnmf_at_baseline_shape(alg):
    """configs[4]: N = M = 4, I = 1025, J = 512, K = 16, incl. the per-(bin, frame) Hermitian eigh of the Wiener
    filter in `separate`; 2 iterations."""
    from oracle import mnmf as omnmf
    from ssspy_b200.bss import FastGaussMNMF
    from ssspy_b200.utils.synth import make_batch
    B, N, I, J, K, n_iter = 2, 4, 1025, 512, 16, 2
    X = make_batch(B, N, I, J, config_id=50, mode="mix")
    m = FastGaussMNMF(n_basis=K, diagonalizer_algorithm=alg, rng=np.random.default_rng(77))
    m.chunk_size = 1
    Y = m(X, n_iter=n_iter)
    rng = np.random.default_rng(77)
    loss = np.asarray(m.loss)
    for b in range(B):
        T = rng.random((N, I, K))
        V = rng.random((N, K, J))
        D = rng.random((I, N, N))
        Q = np.tile(np.eye(N, dtype=np.complex128), (I, 1, 1))
        st = omnmf.run(X[b], T, V, Q, D, n_iter, algorithm=alg)
        ey = relerr(Y[b], st["Y"])
        print("FastGaussMNMF-%s mixture %d: relerr Y %.2e loss %.2e" % (
            alg, b, ey, np.max(np.abs(loss[:, b] / np.asarray(st["loss"]) - 1))))
        assert ey < TOL
        _loss_close(loss[:, b], st["loss"])
