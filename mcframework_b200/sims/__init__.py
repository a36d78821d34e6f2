"""Built-in simulations; the same public names as the reference's ``mcframework/sims/__init__.py:5-23``."""
from __future__ import annotations

from . import black_scholes as _bs
from . import pi as _pi
from . import portfolio as _pf

PiEstimationSimulation = _pi.PiEstimationSimulation
PortfolioSimulation = _pf.PortfolioSimulation
BlackScholesSimulation = _bs.BlackScholesSimulation
BlackScholesPathSimulation = _bs.BlackScholesPathSimulation
# module-level helpers the reference also exports (leading underscore kept for drop-in imports)
_european_payoff = _bs._european_payoff
_simulate_gbm_path = _bs._simulate_gbm_path
_american_exercise_lsm = _bs._american_exercise_lsm

__all__ = sorted(
    ["PiEstimationSimulation", "PortfolioSimulation", "BlackScholesSimulation", "BlackScholesPathSimulation",
     "_european_payoff", "_simulate_gbm_path", "_american_exercise_lsm"], key=lambda s: (s.startswith("_"), s))
