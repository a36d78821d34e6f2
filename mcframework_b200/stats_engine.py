"""StatsEngine / StatsContext and the twelve default metrics, computed by the B200 kernels.

Drop-in for the reference's ``mcframework/stats_engine.py`` (same public names, argument meaning,
return shapes and error behaviour; citations below are to that file), but the arithmetic is not
NumPy's: every metric reads a :class:`DeviceSample` -- the sample resident in HBM plus ONE fused
moments pass (``mcf_moments``: count / non-finite census / mean / M2 / M3 / M4 / sum (x-target)^2) --
and percentiles come from radix selection (``mcf_select``), the bootstrap from the slot-compaction
kernels (``mcf_bootstrap_*``).  What stays on the host is scalar algebra on those outputs and SciPy's
scalar quantile functions.  A missing CUDA library or device raises; there is no CPU fallback.

Engine orchestration (skip / error taxonomy) follows ``StatsEngine.compute`` (:579-661).
"""
from __future__ import annotations

import logging
import math
from dataclasses import dataclass, field, replace
from enum import Enum
from typing import Any, Callable, Generic, Iterable, Mapping, Protocol, Sequence, SupportsFloat, TypeAlias, TypeVar

import numpy as np
from scipy.special import erfinv
from scipy.stats import norm

from ._native import DeviceUnavailableError, NativeError
from .utils import autocrit

logger = logging.getLogger(__name__)
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s - %(message)s"))
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)

_PCTS = (5, 25, 50, 75, 95)
_SMALL_PROB_BOUND = 1e-12
_SMALL_VARIANCE_BOUND = 1e-30
_BCA_JACKKNIFE_DENOMINATOR = 6.0


class MissingContextError(ValueError):
    """A metric needs a context field (``target`` / ``eps``) that was not supplied; the engine skips it."""


class InsufficientDataError(ValueError):
    """Not enough (finite) observations for the requested metric."""


class NanPolicy(str, Enum):
    propagate = "propagate"
    omit = "omit"


class CIMethod(str, Enum):
    auto = "auto"
    z = "z"
    t = "t"
    bootstrap = "bootstrap"


class BootstrapMethod(str, Enum):
    percentile = "percentile"
    bca = "bca"


@dataclass(slots=True)
class StatsContext:
    """Configuration shared by all metrics (reference :167-374; same fields, defaults and validation)."""

    n: int
    confidence: float = 0.95
    ci_method: CIMethod = CIMethod.auto
    percentiles: tuple[int, ...] = (5, 25, 50, 75, 95)
    nan_policy: NanPolicy = NanPolicy.propagate
    target: float | None = None
    eps: float | None = None
    ddof: int = 1
    ess: int | None = None
    rng: int | np.random.Generator | None = None
    n_bootstrap: int = 10_000
    bootstrap: BootstrapMethod = BootstrapMethod.percentile

    def __post_init__(self) -> None:
        if not (0.0 < self.confidence < 1.0):
            raise ValueError("confidence must be in (0,1)")
        for p in self.percentiles:
            if p < 0 or p > 100:
                raise ValueError("percentiles must be in [0,100]")
        if self.n_bootstrap <= 0:
            raise ValueError("n_bootstrap must be > 0")
        if self.ddof < 0:
            raise ValueError("ddof must be >= 0")
        if self.eps is not None and self.eps <= 0:
            raise ValueError("eps must be positive")
        if self.ess is not None:
            if self.ess > self.n:
                raise ValueError(f"ess ({self.ess}) cannot exceed n ({self.n})")
            if self.ess <= 0:
                raise ValueError(f"ess must be positive, got {self.ess}")
        if self.n_bootstrap < 100:
            raise ValueError(f"n_bootstrap ({self.n_bootstrap}) should be >= 100 for reliable estimates")

    def with_overrides(self, **changes) -> "StatsContext":
        return replace(self, **changes)

    @property
    def alpha(self) -> float:
        return 1.0 - self.confidence

    def q_bound(self) -> tuple[float, float]:
        a = self.alpha
        return 100.0 * (a / 2), 100.0 * (1 - a / 2)

    def eff_n(self, observed_len: int, finite_count: int | None = None) -> int:
        # precedence: explicit ess > finite count under "omit" > declared n > observed length
        if self.ess is not None:
            return int(self.ess)
        if self.nan_policy == "omit" and finite_count is not None:
            return int(finite_count)
        return int(self.n or observed_len)

    def get_generators(self) -> np.random.Generator:
        if isinstance(self.rng, np.random.Generator):
            return self.rng
        if isinstance(self.rng, (int, np.integer)):
            return np.random.default_rng(int(self.rng))
        return np.random.default_rng()


@dataclass(frozen=True)
class _CIResult:
    confidence: float
    method: str
    low: float
    high: float
    extras: Mapping[str, SupportsFloat] = field(default_factory=dict)

    def as_dict(self) -> dict[str, float | str]:
        label = self.method.value if hasattr(self.method, "value") else str(self.method)
        d: dict[str, float | str] = {"confidence": float(self.confidence), "method": label, "low": float(self.low),
                                     "high": float(self.high)}
        for key, value in (self.extras or {}).items():
            d[key] = float(value)
        return d


@dataclass(frozen=True)
class ComputeResult:
    """Outcome of :meth:`StatsEngine.compute`: values, plus what was skipped or failed and why."""

    metrics: dict[str, Any]
    skipped: list[tuple[str, str]] = field(default_factory=list)
    errors: list[tuple[str, str]] = field(default_factory=list)

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(\n  metrics={list(self.metrics.keys())},\n"
                f"  skipped={[name for name, _ in self.skipped]},\n  errors={[name for name, _ in self.errors]}\n)")

    def successful_metrics(self) -> set[str]:
        return set(self.metrics.keys())


MetricSet: TypeAlias = Iterable["Metric"]


class Metric(Protocol):
    @property
    def name(self) -> str: ...

    def __call__(self, x: np.ndarray, ctx: StatsContext, /) -> Any: ...


_MetricT = TypeVar("_MetricT")


@dataclass(frozen=True)
class FnMetric(Generic[_MetricT]):
    """Adapter: a named metric backed by ``fn(x, ctx)``."""

    name: str
    fn: Callable[[np.ndarray, StatsContext], _MetricT]
    doc: str = ""

    def __call__(self, x: np.ndarray, ctx: StatsContext) -> _MetricT:
        return self.fn(x, ctx)


# ------------------------------------------------------------------------------------------------
# The sample as the kernels see it
# ------------------------------------------------------------------------------------------------
class DeviceSample:
    """A 1-D float64 sample resident in HBM, with its fused moments pass cached per ``target``.

    ``group`` (a ``torch.distributed`` process group) marks the sample as SHARDED: this rank holds
    ``dev`` and the statistical sample is the concatenation over ranks; the device wrappers then
    combine partials with small collectives (see ``_stats_device``).
    """

    __slots__ = ("dev", "host", "group", "n_local", "_summaries", "_finite", "_full")

    def __init__(self, dev, host=None, group=None):
        self.dev = dev
        self.host = host
        self.group = group
        self.n_local = int(dev.numel())
        self._summaries: dict[float, Any] = {}
        self._finite = None
        self._full = None

    @classmethod
    def of(cls, x) -> "DeviceSample":
        if isinstance(x, DeviceSample):
            return x
        from . import _native

        _native.lib()  # fails loudly without the CUDA library / a device
        import torch

        if isinstance(x, torch.Tensor):
            t = x.detach().to(device="cuda", dtype=torch.float64).reshape(-1).contiguous()
            return cls(t, None)
        host = np.ascontiguousarray(np.asarray(x, dtype=float).reshape(-1))
        return cls(torch.from_numpy(host).to("cuda"), host)

    def summary(self, target: float | None = None):
        from . import _stats_device as SD

        key = 0.0 if target is None else float(target)
        if key not in self._summaries:
            self._summaries[key] = SD.moments(self.dev, key, self.group)
        return self._summaries[key]

    def any_summary(self):
        for s in self._summaries.values():
            return s
        return self.summary(None)

    @property
    def size(self) -> int:
        s = self.any_summary()
        return int(s.n + s.n_nonfinite)

    def full(self):
        """All ranks' values on this rank (needed by the bootstrap gather)."""
        if self.group is None:
            return self.dev
        if self._full is None:
            import torch
            import torch.distributed as dist

            world = dist.get_world_size(self.group)
            sizes = [torch.zeros(1, dtype=torch.int64, device=self.dev.device) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([self.n_local], dtype=torch.int64, device=self.dev.device), group=self.group)
            sizes = [int(s.item()) for s in sizes]
            m = max(sizes)
            pad = torch.zeros(m, dtype=torch.float64, device=self.dev.device)
            pad[: self.n_local] = self.dev
            parts = [torch.empty(m, dtype=torch.float64, device=self.dev.device) for _ in range(world)]
            dist.all_gather(parts, pad, group=self.group)
            self._full = torch.cat([p[:k] for p, k in zip(parts, sizes)])
        return self._full

    def finite_full(self):
        """Finite values only, order preserved (``_clean`` with nan_policy="omit", :734-735)."""
        if self._finite is None:
            from . import _stats_device as SD

            s = self.any_summary()
            self._finite = self.full() if s.n_nonfinite == 0 else SD.compact_finite(self.full())
        return self._finite


def _ensure_ctx(ctx: Any, x) -> StatsContext:
    """Accept a StatsContext, a dict, ``None`` or any attribute bag (reference :664-709)."""
    if isinstance(ctx, StatsContext):
        return ctx
    length = x.size if isinstance(x, DeviceSample) else int(np.asarray(x).size)
    if ctx is None:
        return StatsContext(n=length)
    if isinstance(ctx, dict):
        fields = dict(ctx)
    else:
        try:
            fields = dict(vars(ctx))
        except TypeError as exc:
            raise TypeError("ctx must be a StatsContext, dict, None, or an object with attributes") from exc
    fields.setdefault("n", length)
    return StatsContext(**fields)


class _View:
    """What ``_clean`` (:712-742) yields, expressed through the device summary instead of a copy."""

    __slots__ = ("sample", "s", "omit", "size", "finite_count", "tainted")

    def __init__(self, x, ctx, target=None):
        policy = ctx.nan_policy
        if policy not in ("omit", "propagate", "raise"):
            raise ValueError(f"Unknown nan_policy: {policy}")
        self.sample = DeviceSample.of(x)
        self.s = self.sample.summary(target if target is not None else getattr(ctx, "target", None))
        self.omit = policy == "omit"
        self.finite_count = int(self.s.n)
        total = int(self.s.n + self.s.n_nonfinite)
        self.size = self.finite_count if self.omit else total  # arr.size after cleaning
        self.tainted = (not self.omit) and self.s.n_nonfinite > 0  # NaN/inf take part in the arithmetic

    def mean(self) -> float:
        if not self.tainted:
            return float(self.s.mean)
        s = self.s
        if s.n_nan > 0 or (s.n_posinf > 0 and s.n_neginf > 0):
            return math.nan
        return math.inf if s.n_posinf > 0 else -math.inf

    def std(self, ddof: int) -> float:
        if self.tainted:
            return math.nan
        with np.errstate(all="ignore"):
            return float(np.sqrt(np.float64(self.s.m2) / np.float64(max(self.size - ddof, 0))))

    def data(self):
        return self.sample.finite_full() if self.omit else self.sample.full()


def _clean(x: np.ndarray, ctx: Any) -> tuple[np.ndarray, np.ndarray]:
    """Host-side equivalent of the reference helper, kept for API compatibility (custom metrics)."""
    arr = np.asarray(x, dtype=float)
    finite = np.isfinite(arr)
    if ctx.nan_policy == "omit":
        arr = arr[finite]
    elif ctx.nan_policy not in ("omit", "propagate", "raise"):
        raise ValueError(f"Unknown nan_policy: {ctx.nan_policy}")
    return arr, finite


def _effective_sample_size(x, ctx: StatsContext) -> int:
    v = _View(x, ctx)
    return ctx.eff_n(observed_len=v.size, finite_count=v.finite_count)


def _device_metric(fn):
    fn._mcf_device_metric = True  # StatsEngine.compute hands these the shared DeviceSample
    return fn


# ------------------------------------------------------------------------------------------------
# Metrics
# ------------------------------------------------------------------------------------------------
@_device_metric
def mean(x, ctx: StatsContext) -> float | None:
    """Sample mean (:767-796); ``None`` for an empty sample."""
    ctx = _ensure_ctx(ctx, x)
    v = _View(x, ctx)
    return v.mean() if v.size else None


@_device_metric
def std(x, ctx: StatsContext) -> float | None:
    """Sample standard deviation with ``ctx.ddof`` (:799-833); ``None`` when n_eff <= 1."""
    ctx = _ensure_ctx(ctx, x)
    v = _View(x, ctx)
    if ctx.eff_n(observed_len=v.size, finite_count=v.finite_count) <= 1:
        return None
    return v.std(ctx.ddof)


@_device_metric
def percentiles(x, ctx: StatsContext) -> dict[int, float]:
    """``np.percentile`` semantics (:836-866) via device radix selection; bit-exact order statistics."""
    from . import _stats_device as SD

    ctx = _ensure_ctx(ctx, x)
    v = _View(x, ctx)
    if v.size == 0:
        return {p: float("nan") for p in ctx.percentiles}
    vals = SD.percentiles(v.sample.dev, ctx.percentiles, n_valid=v.size, has_nan=v.s.n_nan > 0,
                          omit_nonfinite=v.omit, group=v.sample.group)
    return dict(zip(ctx.percentiles, map(float, vals)))


def _degenerate(v: _View) -> bool:
    # scipy.stats.skew / kurtosis: all values (numerically) identical -> NaN
    m2 = v.s.m2 / v.size
    return m2 <= (np.finfo(np.float64).eps * v.s.mean) ** 2


@_device_metric
def skew(x, ctx: StatsContext) -> float:
    """Unbiased (G1) skewness, ``scipy.stats.skew(bias=False)`` (:869-902); 0.0 for fewer than 3 values."""
    ctx = _ensure_ctx(ctx, x)
    v = _View(x, ctx)
    n = v.size
    if n <= 2:
        return 0.0
    if v.tainted or _degenerate(v):
        return math.nan
    m2, m3 = v.s.m2 / n, v.s.m3 / n
    return float(math.sqrt((n - 1.0) * n) / (n - 2.0) * m3 / m2 ** 1.5)


@_device_metric
def kurtosis(x, ctx: StatsContext) -> float:
    """Unbiased excess kurtosis (G2), ``scipy.stats.kurtosis(fisher=True, bias=False)`` (:905-937)."""
    ctx = _ensure_ctx(ctx, x)
    v = _View(x, ctx)
    n = v.size
    if n <= 3:
        return 0.0
    if v.tainted or _degenerate(v):
        return math.nan
    m2, m4 = v.s.m2 / n, v.s.m4 / n
    g2 = 1.0 / (n - 2) / (n - 3) * ((n ** 2 - 1.0) * m4 / m2 ** 2.0 - 3 * (n - 1) ** 2.0)
    return float((g2 + 3.0) - 3)


@_device_metric
def ci_mean(x, ctx) -> dict[str, float | str]:
    """z/t confidence interval for the mean (:940-1016)."""
    ctx = _ensure_ctx(ctx, x)
    v = _View(x, ctx)
    nan = float("nan")
    if v.size == 0:
        return _CIResult(ctx.confidence, ctx.ci_method, nan, nan).as_dict()
    n_eff = ctx.eff_n(observed_len=v.size, finite_count=v.finite_count)
    if n_eff < 2:
        return _CIResult(ctx.confidence, ctx.ci_method, nan, nan, extras={"se": nan, "crit": nan}).as_dict()
    mu = v.mean()
    s = v.std(getattr(ctx, "ddof", 1))
    se = 0.0 if s == 0.0 else s / np.sqrt(n_eff)
    crit, method = autocrit(ctx.confidence, n_eff, ctx.ci_method)
    return _CIResult(ctx.confidence, method, mu - crit * se, mu + crit * se, extras={"se": se, "crit": crit}).as_dict()


def _bootstrap_means(arr, n_resamples: int, rng: np.random.Generator):
    """Means of ``n_resamples`` resamples drawn with ``rng.integers(0, n, size=(B, n))`` (:1019-1041).

    ``arr`` is a DeviceSample / device tensor / ndarray; the result is a float64 DEVICE tensor.  The
    index matrix never exists: see ``_stats_device.bootstrap_means``.  ``rng`` is advanced exactly as
    NumPy would advance it."""
    from . import _stats_device as SD

    if isinstance(arr, DeviceSample):
        data, group = arr.full(), arr.group
    else:
        data, group = DeviceSample.of(arr).dev, None
    return SD.bootstrap_means(data, int(n_resamples), rng, group=group)


def _bca_bias_correction(m_hat: float, means_dev) -> float:
    """z0 = Phi^-1(#{means < m_hat} / B), proportion clipped away from 0 and 1 (:1044-1081)."""
    from . import _stats_device as SD

    prop = float(SD.count_less(means_dev, m_hat)) / int(means_dev.numel())
    prop = np.clip(prop, _SMALL_PROB_BOUND, 1 - _SMALL_PROB_BOUND)
    return float(np.sqrt(2) * erfinv(2 * prop - 1))


def _bca_acceleration(summary) -> float:
    """Jackknife acceleration (:1084-1130).  With jack_i = (S - x_i)/(n-1) the deviations are
    d_i = -(x_i - mean)/(n-1), so sum d^3 / (6 (sum d^2)^1.5) = -M3 / (6 M2^1.5): no extra pass."""
    n1 = summary.n - 1.0
    num = -summary.m3 / n1 ** 3
    den = _BCA_JACKKNIFE_DENOMINATOR * (summary.m2 / n1 ** 2) ** 1.5 + _SMALL_VARIANCE_BOUND
    return float(num / den)


def _compute_bca_interval(v: _View, means_dev, confidence: float) -> tuple[float, float]:
    from . import _stats_device as SD

    z0 = _bca_bias_correction(v.mean(), means_dev)
    a = _bca_acceleration(v.s)
    z_lo = float(norm.ppf((1 - confidence) / 2))
    z_hi = float(norm.ppf(1 - (1 - confidence) / 2))

    def shifted(z: float) -> float:
        num = z0 + z
        return float(norm.cdf(z0 + num / (1.0 - a * num))) * 100.0

    p_lo, p_hi = np.clip(shifted(z_lo), 0, 100), np.clip(shifted(z_hi), 0, 100)
    low, high = SD.percentiles(means_dev, [p_lo, p_hi], has_nan=bool(v.tainted and v.s.n_nan > 0))
    return float(low), float(high)


@_device_metric
def ci_mean_bootstrap(x, ctx: StatsContext) -> dict[str, float | str] | None:
    """Percentile or BCa bootstrap interval for the mean (:1202-1301)."""
    from . import _stats_device as SD

    ctx = _ensure_ctx(ctx, x)
    v = _View(x, ctx)
    if v.size == 0:
        return None
    gen = ctx.get_generators()
    data = v.data()
    means = SD.bootstrap_means(data, int(ctx.n_bootstrap), gen, group=v.sample.group)
    lo_q, hi_q = ctx.q_bound()
    # a population with NaN/inf under "propagate" poisons the resample means: NumPy's percentile then returns NaN
    poisoned = bool(v.tainted and SD.moments(means).n_nan > 0)
    if ctx.bootstrap == "percentile" or v.size < 3:
        low, high = SD.percentiles(means, [lo_q, hi_q], has_nan=poisoned)
        return _CIResult(ctx.confidence, "bootstrap-percentile", low, high).as_dict()
    if v.tainted:
        # NaN/inf taking part ("propagate") make the jackknife acceleration NaN, the adjusted percentiles with it, and
        # np.percentile rejects them (:1197) -- after the resamples were drawn, so the generator has moved as well
        raise ValueError("Percentiles must be in the range [0, 100]")
    low, high = _compute_bca_interval(v, means, ctx.confidence)
    return _CIResult(ctx.confidence, "bootstrap-bca", low, high).as_dict()


@_device_metric
def ci_mean_chebyshev(x, ctx: StatsContext) -> dict[str, float | str] | None:
    """Distribution-free interval mean +- k s / sqrt(n), k = 1/sqrt(1-confidence) (:1304-1345)."""
    ctx = _ensure_ctx(ctx, x)
    v = _View(x, ctx)
    n_eff = ctx.eff_n(observed_len=v.size, finite_count=v.finite_count)
    if n_eff < 2:
        return None
    mu, s = mean(v.sample, ctx), std(v.sample, ctx)
    if mu is None or s is None:
        return None
    k = 1.0 / np.sqrt(max(_SMALL_VARIANCE_BOUND, 1.0 - ctx.confidence))
    half = k * s / np.sqrt(n_eff)
    return _CIResult(ctx.confidence, "chebyshev", mu - half, mu + half).as_dict()


@_device_metric
def chebyshev_required_n(x, ctx: StatsContext) -> int:
    """Smallest n with Chebyshev half-width <= eps (:1348-1381)."""
    ctx = _ensure_ctx(ctx, x)
    if ctx.eps is None:
        raise MissingContextError("chebyshev_required_n requires ctx.eps")
    if ctx.eps <= 0:
        raise ValueError("ctx.eps must be positive")
    s = std(x, ctx)
    k = 1.0 / np.sqrt(max(_SMALL_VARIANCE_BOUND, 1.0 - ctx.confidence))
    return int(np.ceil((k * s / float(ctx.eps)) ** 2))


def _mean_sq_to_target(x, ctx) -> float:
    v = _View(x, ctx, target=ctx.target)
    if v.tainted:
        return math.nan if v.s.n_nan > 0 else math.inf
    with np.errstate(all="ignore"):
        return float(np.float64(v.s.sq_target) / np.float64(v.size))


@_device_metric
def markov_error_prob(x, ctx: StatsContext) -> float:
    """Markov bound E[(X-target)^2] / eps^2 (:1384-1419)."""
    ctx = _ensure_ctx(ctx, x)
    if ctx.target is None:
        raise MissingContextError("markov_error_prob requires ctx.target")
    if ctx.eps is None:
        raise MissingContextError("markov_error_prob requires ctx.eps")
    if ctx.eps <= 0:
        raise ValueError("ctx.eps must be positive")
    return _mean_sq_to_target(x, ctx) / ctx.eps ** 2


@_device_metric
def bias_to_target(x, ctx: StatsContext) -> float:
    """mean - target (:1422-1444)."""
    ctx = _ensure_ctx(ctx, x)
    if ctx.target is None:
        raise MissingContextError("bias_to_target requires ctx.target")
    m = mean(x, ctx)
    if m is None:
        raise InsufficientDataError("bias_to_target requires non-empty data after cleaning")
    return float(m - ctx.target)


@_device_metric
def mse_to_target(x, ctx: StatsContext) -> float:
    """mean((x - target)^2) (:1447-1472)."""
    ctx = _ensure_ctx(ctx, x)
    if ctx.target is None:
        raise MissingContextError("mse_to_target requires ctx.target")
    return _mean_sq_to_target(x, ctx)


# ------------------------------------------------------------------------------------------------
# Engine
# ------------------------------------------------------------------------------------------------
class StatsEngine:
    """Evaluates a list of metrics over one sample (:539-661).

    The sample is uploaded once (or taken as-is when the caller already holds a :class:`DeviceSample`,
    as ``MonteCarloSimulation.run`` does) and shared by all built-in metrics; user metrics receive the
    array they were given."""

    def __init__(self, metrics: MetricSet):
        self._metrics = list(metrics)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(metrics=[{', '.join(m.name for m in self._metrics)}])"

    def compute(self, x, ctx: StatsContext | None = None, select: Sequence[str] | None = None, **kwargs: Any) -> ComputeResult:
        if ctx is not None:
            ctx = _ensure_ctx(ctx, x)
        else:
            fields = dict(kwargs)
            fields.setdefault("n", x.size if isinstance(x, DeviceSample) else int(np.asarray(x).size))
            ctx = StatsContext(**fields)
        wanted = self._metrics if select is None else [m for m in self._metrics if m.name in set(select)]
        shared: DeviceSample | None = x if isinstance(x, DeviceSample) else None
        host_view = None
        values: dict[str, Any] = {}
        skipped: list[tuple[str, str]] = []
        errors: list[tuple[str, str]] = []
        for m in wanted:
            on_device = getattr(getattr(m, "fn", m), "_mcf_device_metric", False)
            try:
                if on_device:
                    if shared is None:
                        shared = DeviceSample.of(x)
                    res = m(shared, ctx)
                else:
                    if isinstance(x, DeviceSample):
                        if host_view is None:
                            host_view = x.host if x.host is not None else x.full().cpu().numpy()
                        res = m(host_view, ctx)
                    else:
                        res = m(x, ctx)
            except (DeviceUnavailableError, NativeError):
                raise  # a broken device path is never recorded as a "metric error"
            except MissingContextError as exc:
                skipped.append((m.name, str(exc)))
                logger.debug("Skipping metric %s: %s", m.name, exc)
                continue
            except ValueError as exc:
                if "Missing required context keys" in str(exc):
                    skipped.append((m.name, str(exc)))
                    continue
                raise
            except Exception as exc:  # noqa: BLE001 - same taxonomy as the reference: record and go on
                errors.append((m.name, str(exc)))
                logger.exception("Error computing metric %s", m.name)
                continue
            if res is None:
                skipped.append((m.name, "insufficient data"))
                continue
            values[m.name] = res
        return ComputeResult(metrics=values, skipped=skipped, errors=errors)


def build_default_engine(include_dist_free: bool = True, include_target_bounds: bool = True) -> StatsEngine:
    """The twelve-metric engine in the reference's order (:1475-1528)."""
    table = [
        ("mean", mean, "Sample mean"),
        ("std", std, "Sample standard deviation"),
        ("percentiles", percentiles, "Percentiles over the sample"),
        ("skew", skew, "Fisher skewness (unbiased)"),
        ("kurtosis", kurtosis, "Excess kurtosis (unbiased)"),
        ("ci_mean", ci_mean, "z/t CI for the mean"),
        ("ci_mean_bootstrap", ci_mean_bootstrap, "Bootstrap CI for the mean"),
    ]
    if include_dist_free:
        table += [("ci_mean_chebyshev", ci_mean_chebyshev, "Chebyshev bound CI for the mean"),
                  ("chebyshev_required_n", chebyshev_required_n, "Required n under Chebyshev to reach eps")]
    if include_target_bounds:
        table += [("markov_error_prob", markov_error_prob, "Markov bound P(|X-target|>=eps)"),
                  ("bias_to_target", bias_to_target, "Bias relative to target"),
                  ("mse_to_target", mse_to_target, "Mean squared error to target")]
    return StatsEngine([FnMetric(name, fn, doc) for name, fn, doc in table])


DEFAULT_ENGINE = build_default_engine(include_dist_free=True, include_target_bounds=True)

__all__ = [
    "CIMethod", "NanPolicy", "BootstrapMethod", "StatsContext", "ComputeResult", "Metric", "MetricSet", "FnMetric",
    "StatsEngine", "MissingContextError", "InsufficientDataError", "DeviceSample", "mean", "std", "percentiles", "skew",
    "kurtosis", "ci_mean", "ci_mean_bootstrap", "ci_mean_chebyshev", "chebyshev_required_n", "markov_error_prob",
    "bias_to_target", "mse_to_target", "build_default_engine", "DEFAULT_ENGINE",
]
