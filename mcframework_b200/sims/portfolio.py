"""Portfolio terminal wealth -- reference ``mcframework/sims/portfolio.py:16-84``.

``n = int(years / dt)`` daily steps with ``dt = 1/252``; GBM branch ``V0*exp(sum normal((mu -
sigma^2/2) dt, sigma sqrt(dt)))``, arithmetic branch ``V0*exp(sum log1p(normal(mu dt, sigma sqrt(dt))))``.
The normals are NumPy's ziggurat on the reference stream, replayed by ``mcf_gbm_terminal``.
"""
from __future__ import annotations

from typing import Optional

from numpy.random import Generator

from .. import _native as N
from .. import _programs
from ..core import MonteCarloSimulation
from ._hooks import one_replication

__all__ = ["PortfolioSimulation"]


class PortfolioSimulation(MonteCarloSimulation):
    def __init__(self):
        super().__init__("Portfolio Simulation")

    def _replication_program(self, *, initial_value: float = 10_000.0, annual_return: float = 0.07,
                             volatility: float = 0.20, years: int = 10, use_gbm: bool = True, **_ignored):
        dt = 1.0 / 252.0
        n_steps = int(years / dt)  # computed exactly as the reference's Python expression
        if n_steps == 0:  # a horizon below one day: `rng.normal(size=0)` draws nothing, exp(0) * V0 (:81-84)
            return _programs.constant_program(initial_value)
        mode = N.PORTFOLIO_GBM if use_gbm else N.PORTFOLIO_LOG1P
        return _programs.normals_program(n_steps, [(mode, [initial_value, annual_return, volatility])])

    def single_simulation(self, *, initial_value: float = 10_000.0, annual_return: float = 0.07, volatility: float = 0.20,
                          years: int = 10, use_gbm: bool = True, _rng: Optional[Generator] = None, **kwargs) -> float:
        prog = self._replication_program(initial_value=initial_value, annual_return=annual_return, volatility=volatility,
                                         years=years, use_gbm=use_gbm)
        return one_replication(prog, self._rng(_rng, self.rng))


PortfolioSimulation._program_hook_owner = PortfolioSimulation
