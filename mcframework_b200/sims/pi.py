"""Pi estimation by disk-hit counting -- reference ``mcframework/sims/pi.py:16-82``.

The hook is described to the engine as a replication program: per replication ``2*n_points``
uniform doubles ``-1 + 2u`` taken pairwise as (x, y), hits = #{x*x + y*y <= 1}, result
``4*hits/n_points`` (antithetic: ``n_points//2`` draws counted twice, plus one extra point when
``n_points`` is odd).  The draw count is fixed, so the kernel jumps every lane straight to its own
position in the reference stream (``mcf_pi_hits``); hit counts are bit-exact.
"""
from __future__ import annotations

from typing import Optional

from numpy.random import Generator

from .. import _programs
from ..core import MonteCarloSimulation
from ._hooks import one_replication

__all__ = ["PiEstimationSimulation"]


class PiEstimationSimulation(MonteCarloSimulation):
    def __init__(self):
        super().__init__("Pi Estimation")

    def _replication_program(self, n_points: int = 10_000, antithetic: bool = False, **_ignored):
        return _programs.pi_program(n_points, antithetic)

    def single_simulation(self, n_points: int = 10_000, antithetic: bool = False, _rng: Optional[Generator] = None,
                          **kwargs) -> float:
        return one_replication(_programs.pi_program(n_points, antithetic), self._rng(_rng, self.rng))


PiEstimationSimulation._program_hook_owner = PiEstimationSimulation
