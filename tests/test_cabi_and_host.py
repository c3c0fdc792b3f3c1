"""CPU-only tests: the C-ABI library loads and exports every symbol include/ssb.h declares, the
product path fails loudly without a GPU, and the host-side logic (flooring mapping, pair schedules)
is exact."""
import ctypes
import functools
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ssspy_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "ssb.h")).read()
    declared = set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", header))
    declared -= {"ssb_config", "ssb_plan"}
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), "libssb.so does not export " + name
    # the ctypes table covers the header
    missing = declared - set(_lib.SIGNATURES) - {"ssb_last_error"}
    assert not missing, "no ctypes signature for {}".format(missing)
    assert lib.ssb_version() == 100


def test_config_struct_layout_matches_header():
    from ssspy_b200 import _lib
    # 14 int32/float scalars + 128 pair slots + fast_path + model_param + partitioning + 2 IPA fields
    assert ctypes.sizeof(_lib.SsbConfig) == 4 * (14 + 2 * _lib.SSB_MAX_PAIRS + 6)


def test_plan_validation_runs_without_gpu():
    from ssspy_b200 import _lib
    cfg = _lib.SsbConfig()
    cfg.model, cfg.spatial, cfg.n_batch, cfg.n_sources, cfg.n_bins, cfg.n_frames, cfg.n_basis = 0, 0, 1, 9, 4, 4, 2
    cfg.domain, cfg.normalization = 2.0, 1
    plan = ctypes.c_void_p()
    with pytest.raises(_lib.SsbError, match="n_sources=9 unsupported"):
        _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    cfg.n_sources, cfg.source, cfg.domain = 2, 1, 1.0
    with pytest.raises(_lib.SsbError, match="domain parameter should be 2"):
        _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    cfg.source, cfg.domain = 0, 2.0
    _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    nbytes = ctypes.c_size_t(0)
    _lib.call("ssb_plan_workspace_bytes", plan, ctypes.byref(nbytes))
    assert nbytes.value > 0
    with pytest.raises(_lib.SsbError, match="no buffers bound"):
        _lib.call("ssb_update_once", plan, None)
    _lib.call("ssb_plan_destroy", plan)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ssspy_b200.bss import AuxLaplaceIVA, GaussILRMA
    from ssspy_b200.linalg import inv2
    X = np.zeros((2, 5, 6), dtype=complex)
    for fn in (lambda: GaussILRMA(n_basis=2)(X, n_iter=1), lambda: AuxLaplaceIVA()(X, n_iter=1),
               lambda: inv2(np.eye(2)[None])):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ssspy_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f + " imports the oracle"


def test_flooring_mapping():
    from ssspy_b200 import _lib
    from ssspy_b200.special.flooring import add_flooring, identity, max_flooring
    from ssspy_b200.utils.flooring import choose_flooring_fn, flooring_to_enum
    assert flooring_to_enum(functools.partial(max_flooring, eps=1e-10)) == (_lib.FLOOR_MAX, 1e-10)
    assert flooring_to_enum(max_flooring) == (_lib.FLOOR_MAX, 1e-10)
    assert flooring_to_enum(functools.partial(add_flooring, eps=1e-3)) == (_lib.FLOOR_ADD, 1e-3)
    assert flooring_to_enum(identity) == (_lib.FLOOR_NONE, 0.0)
    assert flooring_to_enum(None) == (_lib.FLOOR_NONE, 0.0)
    with pytest.raises(NotImplementedError):
        flooring_to_enum(lambda x: x)

    class M:
        flooring_fn = staticmethod(max_flooring)
    assert choose_flooring_fn("self", method=M()) is max_flooring
    assert choose_flooring_fn(None) is identity
    np.testing.assert_array_equal(max_flooring(np.array([0.0, 1.0])), [1e-10, 1.0])
    np.testing.assert_array_equal(add_flooring(np.array([0.0, 1.0]), eps=0.5), [0.5, 1.5])


def test_pair_selectors_bit_exact():
    """ssspy/utils/select_pair.py:35-44,:72-76; tests/package/utils/test_select_pair.py."""
    from ssspy_b200.utils.select_pair import combination_pair_selector, sequential_pair_selector, wrap_pairs
    assert list(sequential_pair_selector(2)) == [(0, 1), (1, 0)]
    assert list(sequential_pair_selector(4)) == [(0, 1), (1, 2), (2, 3), (3, 0)]
    assert list(sequential_pair_selector(4, sort=True)) == [(0, 1), (1, 2), (2, 3), (0, 3)]
    assert list(sequential_pair_selector(5, step=2)) == [(0, 1), (2, 3), (4, 0)]
    assert list(sequential_pair_selector(3, stop=5)) == [(0, 1), (1, 2), (2, 0), (0, 1), (1, 2)]
    assert list(combination_pair_selector(3)) == [(0, 1), (0, 2), (1, 2)]
    assert all(m < n for m, n in combination_pair_selector(5, sort=True))
    assert wrap_pairs([(-3, -2), (-1, 0)], 3) == [(0, 1), (2, 0)]
    with pytest.raises(IndexError):
        wrap_pairs([(0, 3)], 3)
    from oracle.spatial import sequential_pairs
    for n in range(2, 9):
        assert list(sequential_pair_selector(n)) == sequential_pairs(n)


def test_shard_range_partition():
    from ssspy_b200.parallel import shard_range
    for B in (1, 7, 64, 512):
        for G in (1, 2, 3, 8):
            spans = [shard_range(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[r][1] == spans[r + 1][0] for r in range(G - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_t_and_ggd_constructor_checks():
    from ssspy_b200.bss import GGDILRMA, TILRMA
    with pytest.raises(AssertionError, match="Shape parameter"):
        GGDILRMA(n_basis=2, beta=2.0)
    with pytest.raises(AssertionError, match="Not support ME"):
        GGDILRMA(n_basis=2, beta=1.0, source_algorithm="ME")
    with pytest.raises(ValueError, match="IPA is not supported for t-ILRMA"):
        TILRMA(n_basis=2, dof=3, spatial_algorithm="IPA")
    with pytest.raises(AssertionError, match="domain parameter should be 2"):
        TILRMA(n_basis=2, dof=3, source_algorithm="ME", domain=1)
    assert repr(TILRMA(n_basis=2, dof=3)).startswith("TILRMA(n_basis=2, dof=3, spatial_algorithm=IP")
    assert repr(GGDILRMA(n_basis=2, beta=1.5)).startswith("GGDILRMA(n_basis=2, beta=1.5, spatial_algorithm=IP")


def test_t_and_ggd_plan_validation():
    from ssspy_b200 import _lib
    cfg = _lib.SsbConfig()
    cfg.model, cfg.spatial, cfg.n_batch, cfg.n_sources, cfg.n_bins, cfg.n_frames, cfg.n_basis = 4, 0, 1, 2, 4, 4, 2
    cfg.domain, cfg.normalization, cfg.model_param = 2.0, 1, 0.0
    plan = ctypes.c_void_p()
    with pytest.raises(_lib.SsbError, match="dof must be positive"):
        _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    cfg.model, cfg.model_param = 5, 2.5
    with pytest.raises(_lib.SsbError, match="Shape parameter"):
        _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    cfg.model_param = 1.0
    _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    _lib.call("ssb_plan_destroy", plan)


def test_fdica_constructor_checks_and_plan_validation():
    from ssspy_b200 import _lib
    from ssspy_b200.bss import AuxFDICA, AuxLaplaceFDICA
    with pytest.raises(ValueError, match="Specify contrast function"):
        AuxFDICA()
    with pytest.raises(AssertionError, match="Not support"):
        AuxLaplaceFDICA(spatial_algorithm="ISS")
    with pytest.raises(ValueError, match="Specify 'reference_id'"):
        AuxLaplaceFDICA(reference_id=None)
    m = AuxLaplaceFDICA(spatial_algorithm="IP2")
    assert repr(m).startswith("AuxLaplaceFDICA(spatial_algorithm=IP2, permutation_alignment=True")
    assert list(m.pair_selector(3)) == [(0, 1), (1, 2), (2, 0)]
    cfg = _lib.SsbConfig()
    cfg.model, cfg.spatial, cfg.n_batch, cfg.n_sources, cfg.n_bins, cfg.n_frames = _lib.MODEL_FDICA_LAPLACE, 2, 1, 2, 4, 4
    cfg.domain = 2.0
    plan = ctypes.c_void_p()
    with pytest.raises(_lib.SsbError, match="Not support spatial algorithm"):
        _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    cfg.spatial = 0
    _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    _lib.call("ssb_plan_destroy", plan)


def test_ipa_plan_validation():
    from ssspy_b200 import _lib
    cfg = _lib.SsbConfig()
    cfg.model, cfg.spatial, cfg.n_batch, cfg.n_sources, cfg.n_bins, cfg.n_frames, cfg.n_basis = 4, 4, 1, 2, 4, 4, 2
    cfg.domain, cfg.model_param = 2.0, 3.0
    plan = ctypes.c_void_p()
    with pytest.raises(_lib.SsbError, match="IPA is not supported for t-ILRMA"):
        _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
    cfg.model, cfg.ipa_newton_iter = 0, -1
    with pytest.raises(_lib.SsbError, match="newton_iter"):
        _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU arm: the unmodified reference's update_once from baseline/_ref, $SSSPY_REF or
    /root/reference when importable, else the oracle port, on the host cores): one JSON line with the keys of the bench
    contract; ranks other than 0 exit 0 without output."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--batch", "4", "--frames", "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    if os.path.isdir(os.path.join(root, "baseline", "_ref", "ssspy")) or os.path.isdir("/root/reference/ssspy"):
        assert line["cpu_baseline"]["kind"] == "reference"
    # the port is still there for a box without the reference
    port = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root,
                          env=dict(os.environ, SSSPY_REF="/nonexistent", SSB_BENCH_FORCE_PORT="1"))
    assert port.returncode == 0 and json.loads(port.stdout.strip().splitlines()[-1])["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_chunk_layout_rules(monkeypatch):
    """Engine chunk plans (ssspy_b200/bss/_engine.py): one plan by default; four chunks for a device-resident batch of
    >= 32 small mixtures; eight for host tensors; explicit chunk_size / SSB_CHUNK win; chunks cover the batch exactly."""
    import types
    from ssspy_b200.bss import _engine
    cls = next(v for v in vars(_engine).values() if isinstance(v, type) and hasattr(v, "_chunk_layout"))

    def layout(B, dims=(2, 1025, 512), cpu=False, chunk_size=None):
        fake = types.SimpleNamespace(_dims=lambda: (B,) + dims, _cpu_tensor_io=cpu, chunk_size=chunk_size)
        return cls._chunk_layout(fake)

    monkeypatch.delenv("SSB_CHUNK", raising=False)
    assert layout(1) == [(0, 1)] and layout(31) == [(0, 31)]
    assert layout(64) == [(0, 16), (16, 32), (32, 48), (48, 64)]
    assert layout(35) == [(0, 9), (9, 18), (18, 27), (27, 35)]
    assert layout(64, dims=(8, 2049, 1024)) == [(0, 64)]            # one mixture already fills the GPU for many waves
    assert len(layout(64, cpu=True)) == 8 and len(layout(35, cpu=True)) == 7 and layout(7, cpu=True) == [(0, 7)]
    assert layout(10, chunk_size=4) == [(0, 4), (4, 8), (8, 10)] and layout(10, chunk_size=10) == [(0, 10)]
    monkeypatch.setenv("SSB_CHUNK", "3")
    got = layout(8, chunk_size=100)
    assert got == [(0, 3), (3, 6), (6, 8)]
    for lay in (got, layout(64), layout(35, cpu=True)):
        assert lay[0][0] == 0 and all(a[1] == b[0] for a, b in zip(lay, lay[1:]))


def test_reference_id_wraps_like_numpy_indexing():
    from ssspy_b200.utils.select_pair import wrap_reference_id
    assert [wrap_reference_id(r, 4) for r in (0, 3, -1, -4)] == [0, 3, 3, 0]
    for bad in (4, -5):
        with pytest.raises(IndexError):
            wrap_reference_id(bad, 4)
        with pytest.raises(IndexError):
            np.zeros(4)[bad]


def test_transform_and_linalg_operator_argument_checks_run_without_gpu():
    """Host-side validation of ssspy_b200.transform.stft / istft (scipy.signal's messages where scipy has one) and of the
    standalone linalg operators happens before anything touches the device; the frame count is a host computation
    (ssb_stft_frames) that matches scipy's."""
    from ssspy_b200 import _lib
    from ssspy_b200.linalg import lqpqm2
    from ssspy_b200.transform import istft, stft
    x = np.zeros((2, 1000))
    with pytest.raises(NotImplementedError, match="window"):
        stft(x, window="hamming", nperseg=64)
    with pytest.raises(ValueError, match="noverlap must be less than nperseg"):
        stft(x, nperseg=64, noverlap=64)
    with pytest.raises(ValueError, match="window must have length of nperseg"):
        stft(x, window=np.ones(32), nperseg=64)
    for kw in (dict(nfft=128), dict(boundary="even"), dict(padded=False), dict(return_onesided=False), dict(axis=0),
               dict(scaling="psd"), dict(detrend="constant")):
        with pytest.raises(NotImplementedError):
            stft(x, nperseg=64, **kw)
    with pytest.raises(NotImplementedError, match="complex"):
        stft(x.astype(np.complex128), nperseg=64)
    with pytest.raises(ValueError, match="at least 2d"):
        istft(np.zeros(33, dtype=np.complex128))
    with pytest.raises(ValueError, match="does not match"):
        istft(np.zeros((33, 10), dtype=np.complex128), nperseg=128)
    with pytest.raises(NotImplementedError):
        lqpqm2(np.zeros((1, 2, 2)), np.zeros((1, 2)), np.zeros(1), singular_fn=lambda v: v < 1e-3)
    # frames as scipy.signal.stft counts them: zero extension by nperseg // 2 on both sides, padded to whole hops
    import scipy.signal as ss
    for n_samples, nperseg, hop in ((1000, 64, 16), (4097, 256, 128), (700, 512, 128), (5000, 1024, 256), (513, 512, 512)):
        n = ctypes.c_int(0)
        _lib.call("ssb_stft_frames", n_samples, nperseg, hop, ctypes.byref(n))
        want = ss.stft(np.zeros(n_samples), nperseg=nperseg, noverlap=nperseg - hop)[2].shape[-1]
        assert n.value == want, (n_samples, nperseg, hop, n.value, want)
