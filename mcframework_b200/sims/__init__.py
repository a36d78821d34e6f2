"""Built-in simulations (same exports as the reference's ``mcframework/sims/__init__.py:5-23``)."""
from __future__ import annotations

from .black_scholes import (BlackScholesPathSimulation, BlackScholesSimulation, _american_exercise_lsm,
                            _european_payoff, _simulate_gbm_path)
from .pi import PiEstimationSimulation
from .portfolio import PortfolioSimulation

__all__ = [
    "PiEstimationSimulation",
    "PortfolioSimulation",
    "BlackScholesSimulation",
    "BlackScholesPathSimulation",
    "_european_payoff",
    "_simulate_gbm_path",
    "_american_exercise_lsm",
]
