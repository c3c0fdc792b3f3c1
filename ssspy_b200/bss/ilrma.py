"""GaussILRMA, TILRMA and GGDILRMA on the device (host mirror of ssspy/bss/ilrma.py: ILRMABase :32-579,
GaussILRMA :582-1989, TILRMA :1992-3334, GGDILRMA :3337-4410).  Same constructors, ``__call__``,
``update_once`` / ``update_source_model`` / ``update_spatial_model`` / ``normalize`` / ``compute_loss`` /
``restore_scale`` / ``apply_projection_back`` / ``apply_minimal_distortion_principle`` and attribute names
(``basis``, ``activation``, ``latent``, ``demix_filter``, ``output``, ``loss``); all arithmetic runs in
libssb.so's CUDA kernels.

Covered: spatial_algorithm IP / IP1 / IP2 / ISS / ISS1 / ISS2 / IPA (IPA for GaussILRMA only, as in the reference),
source_algorithm MM / ME, any ``domain`` in (0, 2], ``partitioning`` False / True (latent variable Z),
normalization True / "power" / "projection_back" / False, projection-back and minimal-distortion-principle scale
restoration.  There is no CPU fallback.
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

__all__ = ["GaussILRMA", "TILRMA", "GGDILRMA"]

spatial_algorithms = ["IP", "IP1", "IP2", "ISS", "ISS1", "ISS2", "IPA"]
source_algorithms = ["MM", "ME"]
PROJECTION_BACK_KEYWORDS = ["projection_back", "projection-back", "PB"]
MINIMAL_DISTORTION_PRINCIPLE_KEYWORDS = ["minimal_distortion_principle", "minimal-distortion-principle", "MDP"]

_SPATIAL_ENUM = {"IP": _lib.SPATIAL_IP1, "IP1": _lib.SPATIAL_IP1, "IP2": _lib.SPATIAL_IP2,
                 "ISS": _lib.SPATIAL_ISS1, "ISS1": _lib.SPATIAL_ISS1, "ISS2": _lib.SPATIAL_ISS2,
                 "IPA": _lib.SPATIAL_IPA}


def _not_on_device(what):
    raise NotImplementedError("{} is not implemented on the device yet and there is no CPU fallback.".format(what))


class ILRMABase(DeviceSeparatorMixin, IterativeMethodBase):
    """ssspy/bss/ilrma.py:32-579."""

    def __init__(self, n_basis, partitioning=False, flooring_fn=functools.partial(max_flooring, eps=EPS),
                 callbacks=None, scale_restoration=True, record_loss=True, reference_id=0, rng=None):
        IterativeMethodBase.__init__(self, callbacks=callbacks, record_loss=record_loss)
        self._init_device_state()
        self.n_basis = n_basis
        self.partitioning = partitioning
        self.flooring_fn = identity if flooring_fn is None else flooring_fn
        self.scale_restoration = scale_restoration
        if reference_id is None and scale_restoration:
            raise ValueError("Specify 'reference_id' if scale_restoration=True.")
        self.reference_id = reference_id
        self.rng = np.random.default_rng() if rng is None else rng

    def __repr__(self):
        s = "ILRMA(n_basis={n_basis}, partitioning={partitioning}, scale_restoration={scale_restoration}"
        s += ", record_loss={record_loss}"
        if self.scale_restoration:
            s += ", reference_id={reference_id}"
        return (s + ")").format(**self.__dict__)

    # ---- state initialisation (ilrma.py:151-270) ---------------------------------------------------
    def _batchify(self, name, tail_shape):
        """Bring a user-supplied state entry to its batched device layout ``(B,) + tail_shape``."""
        t = self._dev(name)
        B = self._dims()[0]
        if tuple(t.shape[1:]) != tuple(tail_shape) or t.shape[0] not in (1, B):
            raise ValueError("{} has shape {} but {} is expected.".format(
                name, tuple(t.shape[1:] if not self._batched else t.shape), tuple(tail_shape)))
        if t.shape[0] != B:
            self._state[name] = t.expand(B, *tail_shape).contiguous()
            self._plan_key = None

    def _reset(self, flooring_fn="self", **kwargs):
        assert self.input is not None, "Specify data!"
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        for key, value in kwargs.items():
            setattr(self, key, value)
        B, N, I, J = self._dims()
        self.n_sources, self.n_channels = N, N  # determined case (ilrma.py:180-181)
        self.n_bins, self.n_frames = I, J
        if not (2 <= N <= _lib.SSB_MAX_SOURCES):
            raise NotImplementedError("n_sources={} is outside the supported range 2..{}.".format(N, _lib.SSB_MAX_SOURCES))
        if not self._has("demix_filter"):
            eye = torch.eye(N, dtype=torch.complex64, device=self._dX.device)
            self._state["demix_filter"] = eye.expand(B, I, N, N).contiguous()
        elif self._state["demix_filter"] is not None:
            self._batchify("demix_filter", (I, N, N))
        W = self._dev("demix_filter")
        if W is None:
            self.separate(self.input, demix_filter=None)  # raises like the reference (None @ ndarray)
        Y = torch.empty_like(self._dX)
        self._initial_separate(W, Y)
        self._state["output"] = Y
        self._host_output = None
        self._init_nmf(flooring_fn=flooring_fn, rng=self.rng)
        self._plan_key = None

    def _init_nmf(self, flooring_fn="self", rng=None):
        """T then V drawn on the host from the caller's Generator in the reference's order
        (ilrma.py:256-268), one mixture after the other for batched input."""
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        if rng is None:
            rng = np.random.default_rng()
        B, N, I, J = self._dims()
        K = self.n_basis
        if not (1 <= K <= _lib.SSB_MAX_BASIS):
            raise NotImplementedError("n_basis={} is outside the supported range 1..{}.".format(K, _lib.SSB_MAX_BASIS))
        if self.partitioning:
            self._init_nmf_partitioned(flooring_fn, rng)
            return
        need_T, need_V = not self._has("basis"), not self._has("activation")
        T = np.empty((B, N, I, K)) if need_T else None
        V = np.empty((B, N, K, J)) if need_V else None
        for b in range(B):
            if need_T:
                T[b] = flooring_fn(rng.random((N, I, K)))
            if need_V:
                V[b] = flooring_fn(rng.random((N, K, J)))
        if need_T:
            self._state["basis"] = _device.to_device(T, torch.float32)
        else:
            self._batchify("basis", (N, I, K))
        if need_V:
            self._state["activation"] = _device.to_device(V, torch.float32)
        else:
            self._batchify("activation", (N, K, J))

    def _init_nmf_partitioned(self, flooring_fn, rng):
        """Z[N,K] (unit column sums, then floored), T[I,K], V[K,J] shared by the sources, drawn in this order
        (ilrma.py:219-245)."""
        B, N, I, J = self._dims()
        K = self.n_basis
        need = {name: not self._has(name) for name in ("latent", "basis", "activation")}
        shapes = {"latent": (N, K), "basis": (I, K), "activation": (K, J)}
        host = {name: np.empty((B,) + shapes[name]) for name in shapes if need[name]}
        for b in range(B):
            for name in ("latent", "basis", "activation"):
                if not need[name]:
                    continue
                x = rng.random(shapes[name])
                if name == "latent":
                    x = x / x.sum(axis=0)
                host[name][b] = flooring_fn(x)
        for name in shapes:
            if need[name]:
                self._state[name] = _device.to_device(host[name], torch.float32)
            else:
                self._batchify(name, shapes[name])

    def _state_rank(self, name):
        if self.partitioning and name in ("basis", "activation"):
            return 3
        return super()._state_rank(name)

    @property
    def _plan_slots(self):
        # the latent variable travels in the `variance` slot of ssb_plan_bind (include/ssb.h)
        return ("demix_filter", "output", "basis", "activation", "latent" if self.partitioning else "variance")

    # ---- C-ABI plan ------------------------------------------------------------------------------
    def _plan_config(self):
        B, N, I, J = self._dims()
        cfg = _lib.SsbConfig()
        cfg.model = self._model
        cfg.model_param = self._model_param()
        cfg.spatial = _SPATIAL_ENUM[self.spatial_algorithm]
        cfg.source = _lib.SOURCE_MM if self.source_algorithm == "MM" else _lib.SOURCE_ME
        cfg.n_batch, cfg.n_sources, cfg.n_bins, cfg.n_frames, cfg.n_basis = B, N, I, J, self.n_basis
        cfg.domain = float(self.domain)
        cfg.flooring, cfg.eps = flooring_to_enum(self.flooring_fn)
        norm = self.normalization
        if not norm:
            cfg.normalization = _lib.NORM_NONE
        elif norm is True or norm == "power":
            cfg.normalization = _lib.NORM_POWER
        elif norm == "projection_back":
            cfg.normalization = _lib.NORM_PROJECTION_BACK
        else:
            raise NotImplementedError("Normalization {} is not implemented.".format(norm))
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
        cfg.partitioning = 1 if self.partitioning else 0
        cfg.ipa_normalization = 1 if getattr(self, "lqpqm_normalization", True) else 0
        cfg.ipa_newton_iter = int(getattr(self, "newton_iter", 1))
        return cfg

    # ---- normalisation (ilrma.py:333-514) ----------------------------------------------------------
    def normalize(self, flooring_fn="self"):
        normalization = self.normalization
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        assert normalization, "Set normalization."
        if type(normalization) is bool:
            normalization = "power"
        if normalization == "power":
            self.normalize_by_power(flooring_fn=flooring_fn)
        elif normalization == "projection_back":
            self.normalize_by_projection_back()
        else:
            raise NotImplementedError("Normalization {} is not implemented.".format(normalization))

    def normalize_by_power(self, flooring_fn="self"):
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        self._with_normalization("power", flooring_fn)

    def normalize_by_projection_back(self):
        if self.partitioning:  # ilrma.py:466-470
            raise NotImplementedError("Projection-back-based normalization is not applicable with partitioning function.")
        self._with_normalization("projection_back", self.flooring_fn)

    def _with_normalization(self, kind, flooring_fn):
        saved = self.normalization
        self.normalization = kind
        try:
            self._set_flooring(flooring_fn)
            self._plan_call("ssb_normalize")
        finally:
            self.normalization = saved

    # ---- loss / scale ----------------------------------------------------------------------------
    def compute_loss(self):
        raise NotImplementedError("Implement 'compute_loss' method.")

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
        """W-modes: W <- projection_back(W), Y <- W X (ilrma.py:557-565); ISS modes:
        Y <- projection_back(Y, X) (ilrma.py:1971-1977)."""
        assert self.scale_restoration, "Set self.scale_restoration=True."
        self._plan_call("ssb_restore_scale")

    def apply_minimal_distortion_principle(self):
        """Y <- mdp(Y, X); W modes refit W = Y X^H (X X^H)^-1 (ssspy/bss/ilrma.py:567-579, iva.py:269-281)."""
        assert self.scale_restoration, "Set self.scale_restoration=True."
        self._plan_call("ssb_restore_scale_mdp")


class _DeviceILRMA(ILRMABase):
    """Everything GaussILRMA (ilrma.py:582), TILRMA (:1992) and GGDILRMA (:3337) share on the device: the three
    differ only in the elementwise source-model factors, which libssb.so selects from ``cfg.model`` /
    ``cfg.model_param``."""

    _model = _lib.MODEL_ILRMA_GAUSS

    def _model_param(self):
        return 0.0

    def _init_algorithms(self, spatial_algorithm, source_algorithm, domain, partitioning, normalization, pair_selector):
        if spatial_algorithm not in _SPATIAL_ENUM:
            _not_on_device("spatial_algorithm={!r}".format(spatial_algorithm))
        self.spatial_algorithm = spatial_algorithm
        self.source_algorithm = source_algorithm
        self.domain = domain
        self.normalization = normalization
        if pair_selector is None:
            if spatial_algorithm in ["IP2", "ISS2"]:
                self.pair_selector = sequential_pair_selector
        else:
            self.pair_selector = pair_selector

    def _repr_fields(self, name, extra):
        s = name + "(n_basis={n_basis}" + extra + ", spatial_algorithm={spatial_algorithm}"
        s += ", source_algorithm={source_algorithm}, domain={domain}, partitioning={partitioning}"
        s += ", normalization={normalization}, scale_restoration={scale_restoration}, record_loss={record_loss}"
        if self.scale_restoration:
            s += ", reference_id={reference_id}"
        return (s + ")").format(**self.__dict__)

    def __call__(self, input, n_iter=100, initial_call=True, **kwargs):
        """Separate ``input`` of shape (n_channels, n_bins, n_frames) [or (batch, ...)] (ilrma.py:820-855)."""
        self.input = input
        self._defer_ok = self._stock_call()
        try:
            self._reset(flooring_fn=self.flooring_fn, **kwargs)
        finally:
            self._defer_ok = False
        if self._stock_call():
            # base.py:48-77 + scale restoration + output copy as one pipeline per chunk of mixtures
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
        """``n_iter`` x ``update_once`` on the current state without loss recording, callbacks or scale
        restoration (the loop of ssspy/bss/base.py:68-77).  Chunks of the batch run concurrently."""
        self._set_flooring(self.flooring_fn)
        self._run_iterations(int(n_iter), False)

    def _stock_call(self):
        """True when nothing user-defined has to run between iterations: no callbacks, no overridden
        update / loss / scale-restoration methods, projection-back (or no) scale restoration."""
        cls = type(self)
        sr = self.scale_restoration
        return (self.callbacks is None and cls.update_once is _DeviceILRMA.update_once
                and cls.compute_loss is _DeviceILRMA.compute_loss and self._stock_update_methods()
                and cls.restore_scale is ILRMABase.restore_scale
                and cls.apply_projection_back is ILRMABase.apply_projection_back
                and (type(sr) is bool or sr in PROJECTION_BACK_KEYWORDS))

    def _stock_update_methods(self):
        cls = type(self)
        return (cls.update_source_model is _DeviceILRMA.update_source_model
                and cls.update_spatial_model is _DeviceILRMA.update_spatial_model
                and cls.normalize is ILRMABase.normalize)

    def _reset(self, flooring_fn="self", **kwargs):
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        super()._reset(flooring_fn=flooring_fn, **kwargs)
        if self.spatial_algorithm in ["ISS", "ISS1", "ISS2", "IPA"]:
            self.demix_filter = None  # state lives in self.output (ilrma.py:897-898)

    def update_once(self, flooring_fn="self"):
        """One iteration: source model, spatial model, normalisation (ilrma.py:900-922)."""
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        if self._stock_update_methods():
            self._set_flooring(flooring_fn)
            self._plan_call("ssb_update_once")
            return
        self.update_source_model(flooring_fn=flooring_fn)
        self.update_spatial_model(flooring_fn=flooring_fn)
        if self.normalization:
            self.normalize(flooring_fn=flooring_fn)

    def update_source_model(self, flooring_fn="self"):
        """MM / ME updates of basis then activation (ilrma.py:924-1005)."""
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        if self.source_algorithm not in source_algorithms:
            raise ValueError("{}-algorithm-based source model updates are not supported.".format(self.source_algorithm))
        self._set_flooring(flooring_fn)
        self._plan_call("ssb_update_source_model")

    def update_spatial_model(self, flooring_fn="self"):
        """IP1 / IP2 on the demixing filters or ISS1 on the spectrograms (ilrma.py:1403-1438)."""
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        if self.spatial_algorithm not in _SPATIAL_ENUM:
            raise NotImplementedError("Not support {}.".format(self.spatial_algorithm))
        self._set_flooring(flooring_fn)
        self._plan_call("ssb_update_spatial_model")

    # the reference's per-algorithm entry points (ilrma.py:980-1005, :1440-1696) dispatch to the same kernels
    def update_source_model_mm(self, flooring_fn="self"):
        assert self.source_algorithm == "MM", "This separator was built with source_algorithm={}.".format(self.source_algorithm)
        self.update_source_model(flooring_fn=flooring_fn)

    def update_source_model_me(self, flooring_fn="self"):
        if self.domain != 2:
            raise ValueError("Domain parameter is expected 2, but given {}.".format(self.domain))
        assert self.source_algorithm == "ME", "This separator was built with source_algorithm={}.".format(self.source_algorithm)
        self.update_source_model(flooring_fn=flooring_fn)

    def update_spatial_model_ip1(self, flooring_fn="self"):
        assert self.spatial_algorithm in ["IP", "IP1"]
        self.update_spatial_model(flooring_fn=flooring_fn)

    def update_spatial_model_ip2(self, flooring_fn="self"):
        assert self.spatial_algorithm == "IP2"
        self.update_spatial_model(flooring_fn=flooring_fn)

    def update_spatial_model_iss1(self, flooring_fn="self"):
        assert self.spatial_algorithm in ["ISS", "ISS1"]
        self.update_spatial_model(flooring_fn=flooring_fn)

    def update_spatial_model_iss2(self, flooring_fn="self"):
        """Pairwise ISS over ``pair_selector(n_sources)`` (ilrma.py:1698-1811)."""
        assert self.spatial_algorithm == "ISS2"
        self.update_spatial_model(flooring_fn=flooring_fn)

    def update_spatial_model_ipa(self, flooring_fn="self"):
        """Iterative projection with adjustment with ``self.lqpqm_normalization`` / ``self.newton_iter``
        (ilrma.py:1813-1908)."""
        assert self.spatial_algorithm == "IPA"
        self.update_spatial_model(flooring_fn=flooring_fn)

    def compute_loss(self):
        """Negative log-likelihood (ilrma.py:1910-1967): a ``float`` for a single mixture, an array of
        shape (batch,) for batched input."""
        return self._loss_from_device()

    # ---- the reference's finer-grained entry points ------------------------------------------------------------
    reconstruct_nmf = _engine_reconstruct_nmf

    def _source_part(self, part, rule, flooring_fn):
        assert self.source_algorithm == rule, \
            "This separator was built with source_algorithm={}.".format(self.source_algorithm)
        if rule == "ME" and self.domain != 2:
            raise ValueError("Domain parameter is expected 2, but given {}.".format(self.domain))
        self._set_flooring(choose_flooring_fn(flooring_fn, method=self))
        self._plan_call("ssb_update_source_part", part)

    def update_latent_mm(self):
        """Latent variable of the partitioning function, MM rule (ilrma.py:1007-1049)."""
        self._source_part(_lib.PART_LATENT, "MM", "self")

    def update_basis_mm(self, flooring_fn="self"):
        """Basis only, MM rule (ilrma.py:1051-1128)."""
        self._source_part(_lib.PART_BASIS, "MM", flooring_fn)

    def update_activation_mm(self, flooring_fn="self"):
        """Activation only, MM rule (ilrma.py:1130-1204)."""
        self._source_part(_lib.PART_ACTIVATION, "MM", flooring_fn)


class _MESubsteps:
    """ME-rule sub-steps of the source model; the reference defines them for GaussILRMA and TILRMA only."""

    def update_latent_me(self):
        """ilrma.py:1206-1247."""
        self._source_part(_lib.PART_LATENT, "ME", "self")

    def update_basis_me(self, flooring_fn="self"):
        """ilrma.py:1249-1325."""
        self._source_part(_lib.PART_BASIS, "ME", flooring_fn)

    def update_activation_me(self, flooring_fn="self"):
        """ilrma.py:1327-1401."""
        self._source_part(_lib.PART_ACTIVATION, "ME", flooring_fn)


class GaussILRMA(_MESubsteps, _DeviceILRMA):
    """ssspy/bss/ilrma.py:582-1989 (signature :752-772)."""

    _ipa_default_kwargs = {"lqpqm_normalization": True, "newton_iter": 1}  # ilrma.py:749-750
    _default_kwargs = _ipa_default_kwargs

    def __init__(self, n_basis, spatial_algorithm="IP", source_algorithm="MM", domain=2, partitioning=False,
                 flooring_fn=functools.partial(max_flooring, eps=EPS), pair_selector=None, callbacks=None,
                 normalization=True, scale_restoration=True, record_loss=True, reference_id=0, rng=None, **kwargs):
        super().__init__(n_basis=n_basis, partitioning=partitioning, flooring_fn=flooring_fn, callbacks=callbacks,
                         scale_restoration=scale_restoration, record_loss=record_loss, reference_id=reference_id,
                         rng=rng)
        assert spatial_algorithm in spatial_algorithms, "Not support {}.".format(spatial_algorithm)
        assert source_algorithm in source_algorithms, "Not support {}.".format(source_algorithm)
        assert 0 < domain <= 2, "domain parameter should be chosen from [0, 2]."
        if source_algorithm == "ME":
            assert domain == 2, "domain parameter should be 2 when you specify ME algorithm."
        self._init_algorithms(spatial_algorithm, source_algorithm, domain, partitioning, normalization, pair_selector)
        # IPA-only keywords are the only valid extras (ilrma.py:802-818)
        valid_keys = set(self._ipa_default_kwargs) if spatial_algorithm == "IPA" else set()
        invalid_keys = set(kwargs) - valid_keys
        assert invalid_keys == set(), "Invalid keywords {} are given.".format(invalid_keys)
        for key, value in kwargs.items():
            setattr(self, key, value)
        for key in valid_keys:
            if not hasattr(self, key):
                setattr(self, key, self._default_kwargs[key])

    def __repr__(self):
        return self._repr_fields("GaussILRMA", "")


class TILRMA(_MESubsteps, _DeviceILRMA):
    """Student-t ILRMA, ssspy/bss/ilrma.py:1992-3334 (signature :2145-2165): ``dof`` is the degree of freedom nu;
    nu -> inf recovers GaussILRMA.  MM and ME source updates (ilrma.py:2384-2827), IP1 / IP2 / ISS1 with the weight
    1 / (nu/(nu+2) (TV)^(2/p) + 2/(nu+2) |y|^2) (ilrma.py:2863-3145), loss :3252-3312."""

    _model = _lib.MODEL_ILRMA_T

    def __init__(self, n_basis, dof, spatial_algorithm="IP", source_algorithm="MM", domain=2, partitioning=False,
                 flooring_fn=functools.partial(max_flooring, eps=EPS), pair_selector=None, callbacks=None,
                 normalization=True, scale_restoration=True, record_loss=True, reference_id=0, rng=None):
        super().__init__(n_basis=n_basis, partitioning=partitioning, flooring_fn=flooring_fn, callbacks=callbacks,
                         scale_restoration=scale_restoration, record_loss=record_loss, reference_id=reference_id,
                         rng=rng)
        assert spatial_algorithm in spatial_algorithms, "Not support {}.".format(spatial_algorithms)
        assert source_algorithm in source_algorithms, "Not support {}.".format(source_algorithm)
        assert 0 < domain <= 2, "domain parameter should be chosen from [0, 2]."
        if spatial_algorithm == "IPA":
            raise ValueError("IPA is not supported for t-ILRMA.")
        if source_algorithm == "ME":
            assert domain == 2, "domain parameter should be 2 when you specify ME algorithm."
        self.dof = dof
        self._init_algorithms(spatial_algorithm, source_algorithm, domain, partitioning, normalization, pair_selector)

    def _model_param(self):
        return float(self.dof)

    def __repr__(self):
        return self._repr_fields("TILRMA", ", dof={dof}")


class GGDILRMA(_DeviceILRMA):
    """Generalised-Gaussian ILRMA, ssspy/bss/ilrma.py:3337-4410 (signature :3490-3510): ``beta`` in (0, 2) is the
    shape parameter (beta -> 2 is the Gaussian case).  MM source updates only (ilrma.py:3698-3905), IP1 / IP2 / ISS1
    with the weight 1 / ((2/beta) floor(|y|^(2-beta)) (TV)^(beta/p)) (ilrma.py:3941-4222), loss :4329-4388."""

    _model = _lib.MODEL_ILRMA_GGD

    def __init__(self, n_basis, beta, spatial_algorithm="IP", source_algorithm="MM", domain=2, partitioning=False,
                 flooring_fn=functools.partial(max_flooring, eps=EPS), pair_selector=None, callbacks=None,
                 normalization=True, scale_restoration=True, record_loss=True, reference_id=0, rng=None):
        super().__init__(n_basis=n_basis, partitioning=partitioning, flooring_fn=flooring_fn, callbacks=callbacks,
                         scale_restoration=scale_restoration, record_loss=record_loss, reference_id=reference_id,
                         rng=rng)
        assert 0 < beta < 2, "Shape parameter {} shoule be chosen from (0, 2).".format(beta)
        assert spatial_algorithm in spatial_algorithms, "Not support {}.".format(spatial_algorithms)
        assert source_algorithm == "MM", "Not support {}.".format(source_algorithm)
        assert 0 < domain <= 2, "domain parameter should be chosen from [0, 2]."
        if spatial_algorithm == "IPA":
            raise ValueError("IPA is not supported for GGD-ILRMA.")
        self.beta = beta
        self._init_algorithms(spatial_algorithm, source_algorithm, domain, partitioning, normalization, pair_selector)

    def _model_param(self):
        return float(self.beta)

    def __repr__(self):
        return self._repr_fields("GGDILRMA", ", beta={beta}")
