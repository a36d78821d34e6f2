"""mcframework_b200 -- the replication loop and StatsEngine reductions of mcframework on B200 GPUs.

Same public names as the reference package (``mcframework/__init__.py:13-28``).  Importing the package
needs neither a GPU nor the CUDA library; running a built-in simulation or a built-in metric does,
and fails loudly without them (no CPU fallback).
"""
from .core import MonteCarloFramework, MonteCarloSimulation, SimulationResult
from .sims import BlackScholesPathSimulation, BlackScholesSimulation, PiEstimationSimulation, PortfolioSimulation
from .stats_engine import DEFAULT_ENGINE, FnMetric, StatsContext, StatsEngine
from .utils import autocrit, t_crit, z_crit

__all__ = [
    "SimulationResult",
    "MonteCarloSimulation",
    "MonteCarloFramework",
    "PiEstimationSimulation",
    "PortfolioSimulation",
    "BlackScholesSimulation",
    "BlackScholesPathSimulation",
    "StatsEngine",
    "StatsContext",
    "FnMetric",
    "DEFAULT_ENGINE",
    "z_crit",
    "t_crit",
    "autocrit",
]
__version__ = "0.1.0"
