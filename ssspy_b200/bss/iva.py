"""Auxiliary-function IVA on the device (host mirror of ssspy/bss/iva.py: IVABase :48-281,
AuxIVABase :553-641, AuxIVA :1403-2214, AuxLaplaceIVA :2976-3128, AuxGaussIVA :3131-3473).

Covered: spatial_algorithm IP / IP1 / IP2 / ISS / ISS1 / ISS2 / IPA with the Laplace and Gauss contrasts.  The
contrast functions are arbitrary Python callables in the reference; on the device they are an
enum, so the generic ``AuxIVA`` accepts only the two known contrasts (no CPU fallback).
"""
import functools

import numpy as np
import torch

from .. import _lib
from ..special.flooring import EPS, identity, max_flooring
from ..utils.flooring import choose_flooring_fn, flooring_to_enum
from ._engine import no_whitening as _no_whitening
from ..utils.select_pair import sequential_pair_selector, wrap_pairs, wrap_reference_id
from ._engine import DeviceSeparatorMixin
from .base import IterativeMethodBase
from .ilrma import (MINIMAL_DISTORTION_PRINCIPLE_KEYWORDS, PROJECTION_BACK_KEYWORDS, _SPATIAL_ENUM,
                    _not_on_device)

__all__ = ["AuxIVA", "AuxLaplaceIVA", "AuxGaussIVA"]

spatial_algorithms = ["IP", "IP1", "IP2", "ISS", "ISS1", "ISS2", "IPA"]


class IVABase(DeviceSeparatorMixin, IterativeMethodBase):
    """ssspy/bss/iva.py:48-281."""

    def __init__(self, flooring_fn=functools.partial(max_flooring, eps=EPS), callbacks=None, scale_restoration=True,
                 record_loss=True, reference_id=0):
        IterativeMethodBase.__init__(self, callbacks=callbacks, record_loss=record_loss)
        self._init_device_state()
        self.flooring_fn = identity if flooring_fn is None else flooring_fn
        self.scale_restoration = scale_restoration
        if reference_id is None and scale_restoration:
            raise ValueError("Specify 'reference_id' if scale_restoration=True.")
        self.reference_id = reference_id

    def __repr__(self):
        s = "IVA(scale_restoration={scale_restoration}, record_loss={record_loss}"
        if self.scale_restoration:
            s += ", reference_id={reference_id}"
        return (s + ")").format(**self.__dict__)

    def _reset(self, **kwargs):
        """iva.py:138-169."""
        assert self.input is not None, "Specify data!"
        for key, value in kwargs.items():
            setattr(self, key, value)
        B, N, I, J = self._dims()
        self.n_sources, self.n_channels = N, N
        self.n_bins, self.n_frames = I, J
        if not (2 <= N <= _lib.SSB_MAX_SOURCES):
            raise NotImplementedError("n_sources={} is outside the supported range 2..{}.".format(N, _lib.SSB_MAX_SOURCES))
        if not self._has("demix_filter"):
            eye = torch.eye(N, dtype=torch.complex64, device=self._dX.device)
            self._state["demix_filter"] = eye.expand(B, I, N, N).contiguous()
        elif self._state["demix_filter"] is not None:
            t = self._dev("demix_filter")
            if tuple(t.shape[1:]) != (I, N, N) or t.shape[0] not in (1, B):
                raise ValueError("demix_filter has shape {} but {} is expected.".format(tuple(t.shape), (I, N, N)))
            if t.shape[0] != B:
                self._state["demix_filter"] = t.expand(B, I, N, N).contiguous()
        W = self._dev("demix_filter")
        if W is None:
            self.separate(self.input, demix_filter=None)  # raises like the reference
        Y = torch.empty_like(self._dX)
        self._initial_separate(W, Y)
        self._state["output"] = Y
        self._host_output = None
        self._plan_key = None

    def compute_loss(self):
        """sum_n mean_j G(y_jn) - 2 sum_i log|det W_i| (iva.py:200-222, :2177-2192)."""
        return self._loss_from_device()

    def restore_scale(self):
        scale_restoration = self.scale_restoration
        assert scale_restoration, "Set self.scale_restoration=True."
        if type(scale_restoration) is bool:
            scale_restoration = PROJECTION_BACK_KEYWORDS[0]
        if scale_restoration in PROJECTION_BACK_KEYWORDS:
            self.apply_projection_back()
        elif scale_restoration in MINIMAL_DISTORTION_PRINCIPLE_KEYWORDS:
            self.apply_minimal_distortion_principle()
        else:
            raise ValueError("{} is not supported for scale restoration.".format(scale_restoration))

    def apply_projection_back(self):
        """iva.py:259-267 (filters) / :2194-2204 (spectrograms, ISS modes)."""
        assert self.scale_restoration, "Set self.scale_restoration=True."
        self._plan_call("ssb_restore_scale")

    def apply_minimal_distortion_principle(self):
        """Y <- mdp(Y, X); W modes refit W = Y X^H (X X^H)^-1 (ssspy/bss/ilrma.py:567-579, iva.py:269-281)."""
        assert self.scale_restoration, "Set self.scale_restoration=True."
        self._plan_call("ssb_restore_scale_mdp")


class AuxIVABase(IVABase):
    """ssspy/bss/iva.py:553-641."""

    def __init__(self, contrast_fn=None, d_contrast_fn=None, flooring_fn=functools.partial(max_flooring, eps=EPS),
                 callbacks=None, scale_restoration=True, record_loss=True, reference_id=0):
        super().__init__(flooring_fn=flooring_fn, callbacks=callbacks, scale_restoration=scale_restoration,
                         record_loss=record_loss, reference_id=reference_id)
        self.contrast_fn = contrast_fn
        self.d_contrast_fn = d_contrast_fn


class AuxIVA(AuxIVABase):
    """ssspy/bss/iva.py:1403-2214 (signature :1582-1598)."""

    _model = None  # set by the Laplace / Gauss subclasses
    _ipa_default_kwargs = {"lqpqm_normalization": True, "newton_iter": 1}  # iva.py:1579-1580
    _default_kwargs = _ipa_default_kwargs

    def __init__(self, spatial_algorithm="IP", contrast_fn=None, d_contrast_fn=None,
                 flooring_fn=functools.partial(max_flooring, eps=EPS), pair_selector=None, callbacks=None,
                 scale_restoration=True, record_loss=True, reference_id=0, **kwargs):
        super().__init__(contrast_fn=contrast_fn, d_contrast_fn=d_contrast_fn, flooring_fn=flooring_fn,
                         callbacks=callbacks, scale_restoration=scale_restoration, record_loss=record_loss,
                         reference_id=reference_id)
        assert spatial_algorithm in spatial_algorithms, "Not support {}.".format(spatial_algorithm)
        if spatial_algorithm not in _SPATIAL_ENUM:
            _not_on_device("spatial_algorithm={!r}".format(spatial_algorithm))
        self.spatial_algorithm = spatial_algorithm
        if pair_selector is None:
            if spatial_algorithm in ["IP2", "ISS2"]:
                self.pair_selector = sequential_pair_selector
        else:
            self.pair_selector = pair_selector
        # IPA-only keywords are the only valid extras (iva.py:1619-1635)
        valid_keys = set(self._ipa_default_kwargs) if spatial_algorithm == "IPA" else set()
        invalid_keys = set(kwargs) - valid_keys
        assert invalid_keys == set(), "Invalid keywords {} are given.".format(invalid_keys)
        for key, value in kwargs.items():
            setattr(self, key, value)
        for key in valid_keys:
            if not hasattr(self, key):
                setattr(self, key, self._default_kwargs[key])

    def __call__(self, input, n_iter=100, initial_call=True, **kwargs):
        """iva.py:1637-1672."""
        if self._model is None:
            _not_on_device("AuxIVA with user-defined contrast functions (use AuxLaplaceIVA / AuxGaussIVA)")
        self.input = input
        self._defer_ok = self._stock_call()
        try:
            self._reset(**kwargs)
        finally:
            self._defer_ok = False
        if self._stock_call():
            self._stock_pipeline(n_iter, initial_call, pb=bool(self.scale_restoration))
            return self.output
        IterativeMethodBase.__call__(self, n_iter=n_iter, initial_call=initial_call)
        if self.scale_restoration:
            self.restore_scale()
        elif self._state.get("demix_filter") is not None:
            self._plan_call("ssb_plan_separate")
        _lib.check_status()  # LinAlgError where the reference's np.linalg.solve / inv would have raised
        return self.output

    def run_iterations(self, n_iter):
        """``n_iter`` x ``update_once`` on the current state (ssspy/bss/base.py:68-77) without loss
        recording, callbacks or scale restoration."""
        self._set_flooring(self.flooring_fn)
        self._run_iterations(int(n_iter), False)

    def _stock_call(self):
        cls = type(self)
        sr = self.scale_restoration
        return (self.callbacks is None and cls.update_once in (AuxIVA.update_once, AuxGaussIVA.update_once)
                and cls.compute_loss is IVABase.compute_loss
                and cls.update_source_model in (AuxIVA.update_source_model, AuxGaussIVA.update_source_model)
                and cls.restore_scale is IVABase.restore_scale
                and cls.apply_projection_back is IVABase.apply_projection_back
                and (type(sr) is bool or sr in PROJECTION_BACK_KEYWORDS))

    def __repr__(self):
        s = "AuxIVA(spatial_algorithm={spatial_algorithm}, scale_restoration={scale_restoration}"
        s += ", record_loss={record_loss}"
        if self.scale_restoration:
            s += ", reference_id={reference_id}"
        return (s + ")").format(**self.__dict__)

    def _reset(self, **kwargs):
        super()._reset(**kwargs)
        if self.spatial_algorithm in ["ISS", "ISS1", "ISS2", "IPA"]:
            self.demix_filter = None  # iva.py:1696-1697

    def _plan_config(self):
        B, N, I, J = self._dims()
        cfg = _lib.SsbConfig()
        cfg.model = self._model
        cfg.spatial = _SPATIAL_ENUM[self.spatial_algorithm]
        cfg.source = _lib.SOURCE_MM
        cfg.n_batch, cfg.n_sources, cfg.n_bins, cfg.n_frames, cfg.n_basis = B, N, I, J, 0
        cfg.domain = 2.0
        cfg.flooring, cfg.eps = flooring_to_enum(self.flooring_fn)
        cfg.normalization = _lib.NORM_NONE
        cfg.reference_id = 0 if self.reference_id is None else wrap_reference_id(self.reference_id, N)
        pairs = []
        if cfg.spatial in (_lib.SPATIAL_IP2, _lib.SPATIAL_ISS2):
            pairs = wrap_pairs(self.pair_selector(N), N)
            if len(pairs) > _lib.SSB_MAX_PAIRS:
                raise NotImplementedError("more than {} pairs per iteration".format(_lib.SSB_MAX_PAIRS))
        cfg.n_pairs = len(pairs)
        for q, (m, n) in enumerate(pairs):
            cfg.pairs[2 * q], cfg.pairs[2 * q + 1] = m, n
        cfg.fast_path = 1 if getattr(self, "fast_path", True) else 0
        cfg.no_whitening = _no_whitening(self)
        cfg.ipa_normalization = 1 if getattr(self, "lqpqm_normalization", True) else 0
        cfg.ipa_newton_iter = int(getattr(self, "newton_iter", 1))
        return cfg

    def update_source_model(self):
        """No source parameters for the Laplace contrast; AuxGaussIVA overrides (iva.py:3465-3473)."""

    def update_once(self, flooring_fn="self"):
        """Auxiliary weights from the current separation, then IP1 / IP2 / ISS1 (iva.py:1699-1734)."""
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        if self.spatial_algorithm not in _SPATIAL_ENUM:
            raise NotImplementedError("Not support {}.".format(self.spatial_algorithm))
        self._set_flooring(flooring_fn)
        self._plan_call("ssb_update_spatial_model")

    def update_once_ip1(self, flooring_fn="self"):
        assert self.spatial_algorithm in ["IP", "IP1"]
        AuxIVA.update_once(self, flooring_fn=flooring_fn)

    def update_once_ip2(self, flooring_fn="self"):
        assert self.spatial_algorithm == "IP2"
        AuxIVA.update_once(self, flooring_fn=flooring_fn)

    def update_once_iss1(self, flooring_fn="self"):
        assert self.spatial_algorithm in ["ISS", "ISS1"]
        AuxIVA.update_once(self, flooring_fn=flooring_fn)

    def update_once_iss2(self, flooring_fn="self"):
        """Pairwise ISS over ``pair_selector(n_sources)`` with weights from the current output (iva.py:1968-2066)."""
        assert self.spatial_algorithm == "ISS2"
        AuxIVA.update_once(self, flooring_fn=flooring_fn)

    def update_once_ipa(self, flooring_fn="self"):
        """Iterative projection with adjustment with weights from the current output (iva.py:2068-2176)."""
        assert self.spatial_algorithm == "IPA"
        AuxIVA.update_once(self, flooring_fn=flooring_fn)


class AuxLaplaceIVA(AuxIVA):
    """Spherical Laplace contrast G(r) = 2 r, G'(r) = 2 (ssspy/bss/iva.py:2976-3128)."""

    _model = _lib.MODEL_IVA_LAPLACE

    def __init__(self, spatial_algorithm="IP", flooring_fn=functools.partial(max_flooring, eps=EPS),
                 pair_selector=None, callbacks=None, scale_restoration=True, record_loss=True, reference_id=0,
                 **kwargs):
        def contrast_fn(y):
            return 2 * np.linalg.norm(y, axis=1)

        def d_contrast_fn(y):
            return 2 * np.ones_like(y)

        super().__init__(spatial_algorithm=spatial_algorithm, contrast_fn=contrast_fn, d_contrast_fn=d_contrast_fn,
                         flooring_fn=flooring_fn, pair_selector=pair_selector, callbacks=callbacks,
                         scale_restoration=scale_restoration, record_loss=record_loss, reference_id=reference_id,
                         **kwargs)


class AuxGaussIVA(AuxIVA):
    """Time-varying Gaussian contrast (ssspy/bss/iva.py:3131-3473): ``variance[n,j] = mean_i |y|^2``,
    G = I log(alpha) + r^2 / alpha, G' = 2 r / alpha."""

    _model = _lib.MODEL_IVA_GAUSS

    def __init__(self, spatial_algorithm="IP", flooring_fn=functools.partial(max_flooring, eps=EPS),
                 pair_selector=None, callbacks=None, scale_restoration=True, record_loss=True, reference_id=0,
                 **kwargs):
        def contrast_fn(y):
            alpha = self.variance
            norm = np.linalg.norm(y, axis=1)
            return self.n_bins * np.log(alpha) + (norm ** 2) / alpha

        def d_contrast_fn(y, variance=None):
            alpha = self.variance if variance is None else variance
            return 2 * y / alpha

        super().__init__(spatial_algorithm=spatial_algorithm, contrast_fn=contrast_fn, d_contrast_fn=d_contrast_fn,
                         flooring_fn=flooring_fn, pair_selector=pair_selector, callbacks=callbacks,
                         scale_restoration=scale_restoration, record_loss=record_loss, reference_id=reference_id,
                         **kwargs)

    def _reset(self, **kwargs):
        super()._reset(**kwargs)
        B, N, I, J = self._dims()
        self._state["variance"] = torch.ones((B, N, J), dtype=torch.float32, device=self._dX.device)  # iva.py:3317
        self._plan_key = None

    def update_once(self, flooring_fn="self"):
        """iva.py:3319-3337."""
        self.update_source_model()
        super().update_once(flooring_fn=flooring_fn)

    def update_source_model(self):
        """variance = mean_i |y|^2 from the current separation, no floor (iva.py:3465-3473)."""
        self._plan_call("ssb_update_source_model")
