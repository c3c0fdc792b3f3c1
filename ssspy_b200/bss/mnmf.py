"""FastGaussMNMF on the device (host mirror of ssspy/bss/mnmf.py: MNMFBase :21-297, FastMNMFBase
:417-678, FastGaussMNMF :1076-1675; BASELINE.json config 5 maps to this class, the reference has no
"FastMNMF2").  Same constructor and methods; ``partitioning`` is unsupported in the reference itself.

Covered: ``diagonalizer_algorithm`` IP / IP1 / IP2, power normalisation, the multichannel Wiener filter
``separate`` with its per-(bin, frame) Hermitian eigendecomposition (``to_psd``), loss, state injection
(``basis``, ``activation``, ``spatial``, ``diagonalizer``), batched input.  Only the determined case
``n_sources == n_channels`` runs on the device.  ``instant_covariance`` is not materialised: the
reference computes it in ``_reset`` (I*J eigendecompositions) but FastGaussMNMF never reads it
(SURVEY.md Appendix A.6).
"""
import functools

import numpy as np
import torch

from .. import _device, _lib
from ..special.flooring import EPS, identity, max_flooring
from ..utils.flooring import choose_flooring_fn, flooring_to_enum
from ._engine import no_whitening as _no_whitening
from ..utils.select_pair import sequential_pair_selector, wrap_pairs, wrap_reference_id
from ._engine import DeviceSeparatorMixin
from ._engine import reconstruct_nmf as _engine_reconstruct_nmf
from .base import IterativeMethodBase
from .ilrma import _not_on_device

__all__ = ["FastGaussMNMF"]

diagonalizer_algorithms = ["IP", "IP1", "IP2"]


class MNMFBase(DeviceSeparatorMixin, IterativeMethodBase):
    """ssspy/bss/mnmf.py:21-297."""

    _plan_slots = ("diagonalizer", "output", "basis", "activation", "spatial")

    def __init__(self, n_basis, n_sources=None, partitioning=False, flooring_fn=functools.partial(max_flooring, eps=EPS),
                 callbacks=None, normalization=True, record_loss=True, reference_id=0, rng=None):
        IterativeMethodBase.__init__(self, callbacks=callbacks, record_loss=record_loss)
        self._init_device_state()
        self.n_basis = n_basis
        self.n_sources = n_sources
        self.partitioning = partitioning
        self.flooring_fn = identity if flooring_fn is None else flooring_fn
        self.normalization = normalization
        self.reference_id = reference_id
        self.rng = np.random.default_rng() if rng is None else rng

    def __call__(self, input, n_iter=100, initial_call=True, **kwargs):
        """mnmf.py MNMFBase.__call__: reset, iterate, ``output = separate(input)``."""
        self.input = input
        self._reset(**kwargs)
        cls = type(self)
        stock = (self.callbacks is None and cls.update_once is FastGaussMNMF.update_once
                 and cls.compute_loss is FastGaussMNMF.compute_loss and cls.separate is FastGaussMNMF.separate)
        if stock:
            self._set_flooring(self.flooring_fn)
            rec = bool(self.record_loss)
            losses = self._run_iterations(n_iter, rec, initial_loss=bool(initial_call and rec),
                                          tail=lambda ch, sp: _lib.call("ssb_plan_separate", ch["plan"], sp))
            if losses is not None:
                self.loss.extend(losses[i].copy() if self._batched else float(losses[i, 0])
                                 for i in range(losses.shape[0]))
        else:
            IterativeMethodBase.__call__(self, n_iter=n_iter, initial_call=initial_call)
            self._plan_call("ssb_plan_separate")
        _lib.check_status()  # LinAlgError where the reference's np.linalg.solve / inv would have raised
        return self.output


class FastGaussMNMF(MNMFBase):
    """ssspy/bss/mnmf.py:1076-1675."""

    def __init__(self, n_basis, n_sources=None, diagonalizer_algorithm="IP", partitioning=False,
                 flooring_fn=functools.partial(max_flooring, eps=EPS), pair_selector=None, callbacks=None,
                 normalization=True, record_loss=True, reference_id=0, rng=None):
        super().__init__(n_basis, n_sources=n_sources, partitioning=partitioning, flooring_fn=flooring_fn,
                         callbacks=callbacks, normalization=normalization, record_loss=record_loss,
                         reference_id=reference_id, rng=rng)
        assert diagonalizer_algorithm in diagonalizer_algorithms, "Not support {}.".format(diagonalizer_algorithm)
        assert not partitioning, "partitioning function is not supported."
        self.diagonalizer_algorithm = diagonalizer_algorithm
        if pair_selector is None:
            if diagonalizer_algorithm == "IP2":
                self.pair_selector = sequential_pair_selector
        else:
            self.pair_selector = pair_selector

    def __repr__(self):
        s = "FastGaussMNMF(n_basis={n_basis}"
        if self.n_sources is not None:
            s += ", n_sources={n_sources}"
        if hasattr(self, "n_channels"):
            s += ", n_channels={n_channels}"
        s += ", diagonalizer_algorithm={diagonalizer_algorithm}, partitioning={partitioning}"
        s += ", record_loss={record_loss}, reference_id={reference_id})"
        return s.format(**self.__dict__)

    # ---- state initialisation (mnmf.py:499-540, :557-596) ------------------------------------------------
    def _fit(self, name, tail_shape):
        t = self._dev(name)
        B = self._dims()[0]
        if tuple(t.shape[1:]) != tuple(tail_shape) or t.shape[0] not in (1, B):
            raise ValueError("{} has shape {} but {} is expected.".format(name, tuple(t.shape), tuple(tail_shape)))
        if t.shape[0] != B:
            self._state[name] = t.expand(B, *tail_shape).contiguous()

    def _reset(self, flooring_fn="self", **kwargs):
        assert self.input is not None, "Specify data!"
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        for key, value in kwargs.items():
            setattr(self, key, value)
        B, M, I, J = self._dims()
        N = M if self.n_sources is None else self.n_sources
        if N != M:
            _not_on_device("FastGaussMNMF with n_sources != n_channels")
        if not (2 <= N <= _lib.SSB_MAX_SOURCES):
            raise NotImplementedError("n_sources={} is outside the supported range 2..{}.".format(N, _lib.SSB_MAX_SOURCES))
        self.n_sources, self.n_channels = N, M
        self.n_bins, self.n_frames = I, J
        K = self.n_basis
        rng = self.rng
        # draw order of the reference: T, V (mnmf.py:251-263), then D (mnmf.py:594-596); Q = identity
        need_T, need_V, need_D = not self._has("basis"), not self._has("activation"), not self._has("spatial")
        T = np.empty((B, N, I, K)) if need_T else None
        V = np.empty((B, N, K, J)) if need_V else None
        D = np.empty((B, I, N, M)) if need_D else None
        for b in range(B):  # one mixture after the other, as running the reference on each would draw
            if need_T:
                T[b] = flooring_fn(rng.random((N, I, K)))
            if need_V:
                V[b] = flooring_fn(rng.random((N, K, J)))
            if need_D:
                D[b] = flooring_fn(rng.random((I, N, M)))
        for name, arr, shape in (("basis", T, (N, I, K)), ("activation", V, (N, K, J)), ("spatial", D, (I, N, M))):
            if arr is not None:
                self._state[name] = _device.to_device(arr, torch.float32)
            else:
                self._fit(name, shape)
        if not self._has("diagonalizer"):
            eye = torch.eye(M, dtype=torch.complex64, device=self._dX.device)
            self._state["diagonalizer"] = eye.expand(B, I, M, M).contiguous()
        else:
            self._fit("diagonalizer", (I, M, M))
        if self._pending_h2d is not None:
            self._dX.copy_(self._pending_h2d, non_blocking=True)
            self._pending_h2d = None
        self._state["output"] = torch.empty_like(self._dX)
        self._host_output = None
        self._plan_key = None
        self._plan_call("ssb_plan_separate")  # self.output = self.separate(X)  (mnmf.py:540)

    def _plan_config(self):
        B, N, I, J = self._dims()
        cfg = _lib.SsbConfig()
        cfg.model = _lib.MODEL_FASTMNMF_GAUSS
        cfg.spatial = _lib.SPATIAL_IP2 if self.diagonalizer_algorithm == "IP2" else _lib.SPATIAL_IP1
        cfg.source = _lib.SOURCE_MM
        cfg.n_batch, cfg.n_sources, cfg.n_bins, cfg.n_frames, cfg.n_basis = B, N, I, J, self.n_basis
        cfg.domain = 2.0
        cfg.flooring, cfg.eps = flooring_to_enum(self.flooring_fn)
        norm = self.normalization
        if not norm:
            cfg.normalization = _lib.NORM_NONE
        elif norm is True or norm == "power":
            cfg.normalization = _lib.NORM_POWER
        else:
            raise NotImplementedError("Normalization {} is not implemented.".format(norm))
        cfg.reference_id = wrap_reference_id(self.reference_id, N)
        pairs = wrap_pairs(self.pair_selector(N), N) if cfg.spatial == _lib.SPATIAL_IP2 else []
        if len(pairs) > _lib.SSB_MAX_PAIRS:
            raise NotImplementedError("more than {} pairs per iteration".format(_lib.SSB_MAX_PAIRS))
        cfg.n_pairs = len(pairs)
        for q, (m, n) in enumerate(pairs):
            cfg.pairs[2 * q], cfg.pairs[2 * q + 1] = m, n
        cfg.fast_path = 1
        cfg.no_whitening = _no_whitening(self)
        return cfg

    # ---- the reference's methods ----------------------------------------------------------------------
    def separate(self, input):
        """Multichannel Wiener filter of the current model applied to ``input`` (mnmf.py:1174-1217).  The
        device plan is bound to ``self.input``; other inputs are not supported on the device."""
        if input is not self.input and not (_device.is_tensor(input) and input.data_ptr() == self._dX.data_ptr()):
            same = (not _device.is_tensor(input)) and self._input_host is not None and \
                np.shares_memory(input, self._input_host)
            if not same and not np.array_equal(np.asarray(input), np.asarray(self.input)):
                _not_on_device("FastGaussMNMF.separate on an input other than the one the separator was called with")
        self._plan_call("ssb_plan_separate")
        return self.output

    def compute_loss(self):
        """mnmf.py:1219-1261."""
        return self._loss_from_device()

    def update_once(self, flooring_fn="self"):
        """basis, activation, diagonaliser, spatial, normalisation (mnmf.py:1278-1303)."""
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        cls = type(self)
        if all(getattr(cls, m) is getattr(FastGaussMNMF, m) for m in
               ("update_basis", "update_activation", "update_diagonalizer", "update_spatial", "normalize")):
            self._set_flooring(flooring_fn)
            self._plan_call("ssb_update_once")
            return
        self.update_basis(flooring_fn=flooring_fn)
        self.update_activation(flooring_fn=flooring_fn)
        self.update_diagonalizer(flooring_fn=flooring_fn)
        self.update_spatial()
        if self.normalization:
            self.normalize(flooring_fn=flooring_fn)

    def run_iterations(self, n_iter):
        """``n_iter`` x ``update_once`` on the current state without loss recording or callbacks (the loop of
        ssspy/bss/base.py:68-77) as one ``ssb_run`` per chunk plan: inside it the spatial sweep hands Z2 = |Q x|^2 of the
        new diagonaliser to the source model of the next iteration."""
        self._set_flooring(self.flooring_fn)
        self._run_iterations(int(n_iter), False)

    def update_source_model(self, flooring_fn="self"):
        """``update_basis`` then ``update_activation`` (mnmf.py:1305-1417)."""
        self._set_flooring(choose_flooring_fn(flooring_fn, method=self))
        self._plan_call("ssb_update_source_model")

    def update_spatial_model(self, flooring_fn="self"):
        """``update_diagonalizer`` then ``update_spatial`` (mnmf.py:1419-1675)."""
        self._set_flooring(choose_flooring_fn(flooring_fn, method=self))
        self._plan_call("ssb_update_spatial_model")

    def update_basis(self, flooring_fn="self"):
        _not_on_device("FastGaussMNMF.update_basis on its own (use update_source_model)")

    def update_activation(self, flooring_fn="self"):
        _not_on_device("FastGaussMNMF.update_activation on its own (use update_source_model)")

    def update_diagonalizer(self, flooring_fn="self"):
        _not_on_device("FastGaussMNMF.update_diagonalizer on its own (use update_spatial_model)")

    def update_spatial(self):
        _not_on_device("FastGaussMNMF.update_spatial on its own (use update_spatial_model)")

    def update_diagonalizer_ip1(self, flooring_fn="self"):
        _not_on_device("FastGaussMNMF.update_diagonalizer_ip1 on its own (use update_spatial_model)")

    def update_diagonalizer_ip2(self, flooring_fn="self"):
        _not_on_device("FastGaussMNMF.update_diagonalizer_ip2 on its own (use update_spatial_model)")

    def normalize_by_power(self, flooring_fn="self"):
        """mnmf.py:632-678."""
        self._set_flooring(choose_flooring_fn(flooring_fn, method=self))
        self._plan_call("ssb_normalize")

    reconstruct_nmf = _engine_reconstruct_nmf

    def normalize(self, flooring_fn="self"):
        """mnmf.py:632-678 (power normalisation of Q and D)."""
        normalization = self.normalization
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        assert normalization, "Set normalization."
        if type(normalization) is bool:
            normalization = "power"
        if normalization != "power":
            raise NotImplementedError("Normalization {} is not implemented.".format(normalization))
        self._set_flooring(flooring_fn)
        self._plan_call("ssb_normalize")
