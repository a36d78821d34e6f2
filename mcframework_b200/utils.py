"""Critical values for confidence intervals (host-side scalars; SciPy's quantile functions).

Mirrors the reference's ``mcframework/utils.py`` (``z_crit`` :35-64, ``t_crit`` :67-97, ``autocrit``
:100-165): same names, argument meaning and error messages.  These are scalar evaluations of the
normal / Student-t quantile function -- nothing here belongs on a GPU.
"""
from __future__ import annotations

from scipy.stats import norm as _norm
from scipy.stats import t as _student

__all__ = ["z_crit", "t_crit", "autocrit"]


def _check_level(confidence: float) -> None:
    if not (0.0 < confidence < 1.0):
        raise ValueError("confidence must be in the open interval (0, 1)")


def _upper_tail_point(confidence: float) -> float:
    return 1.0 - (1.0 - confidence) / 2.0


def z_crit(confidence: float) -> float:
    """Two-sided standard-normal critical value: ``z`` with ``P(|Z| <= z) = confidence``."""
    _check_level(confidence)
    return float(_norm.ppf(_upper_tail_point(confidence)))


def t_crit(confidence: float, df: int) -> float:
    """Two-sided Student-t critical value with ``df`` degrees of freedom."""
    _check_level(confidence)
    if df < 1:
        raise ValueError("df must be >= 1")
    return float(_student.ppf(_upper_tail_point(confidence), df))


def autocrit(confidence: float, n: int, method: str = "auto") -> tuple[float, str]:
    """``(critical value, "z" | "t")``; ``"auto"`` switches from t to z at ``n >= 30``."""
    _check_level(confidence)
    if method not in ("auto", "z", "t"):
        raise ValueError("method must be 'auto', 'z', or 't'")
    use_z = method == "z" or (method == "auto" and n >= 30)
    if use_z:
        return z_crit(confidence), "z"
    return t_crit(confidence, max(1, n - 1)), "t"
