"""Declared draw programs for simple user simulations.

A subclass of ``MonteCarloSimulation`` whose hook is a few ``Generator`` calls followed by a reduction can SAY so, and
its replications then run as a kernel on the reference streams instead of a Python loop::

    class Dice(MonteCarloSimulation):
        def single_simulation(self, _rng=None, **kw):            # the definition (and what a CPU user runs)
            return float(self._rng(_rng, self.rng).integers(1, 7, size=2).sum())

        def draw_program(self, **kw):                            # the same thing, declared
            return programs.integers_sum(1, 7, size=2)

The declaration is a promise that both agree; results are bit-identical to the reference at matched seeds (integer
draws follow NumPy's buffered 32-bit Lemire rule; a rejected slot -- probability (2**32 mod span)/2**32 per draw -- makes
the engine run the Python hook instead, so the promise also holds in that case).

=====================  =========================================================================
``integers_sum``        ``float(rng.integers(low, high, size=size).sum())``        (reference core.py:22-25)
``integers_longest_run`` longest run of ``value`` in ``rng.integers(low, high, size=size)``
                         (docs/source/getting-started.md:139-183, the coin-flip streak)
``normal``              ``float(rng.normal(loc, scale))``                          (reference tests/conftest.py:9-14)
=====================  =========================================================================
"""
from __future__ import annotations

from . import _native as N
from . import _programs

__all__ = ["integers_sum", "integers_longest_run", "normal"]


def integers_sum(low: int, high: int, size: int):
    return _programs.integers_program(low, high, size, op=0)


def integers_longest_run(low: int, high: int, size: int, value: int = 1):
    return _programs.integers_program(low, high, size, op=1, match=value)


def normal(loc: float = 0.0, scale: float = 1.0):
    return _programs.normals_program(1, [(N.NORMAL_LOC_SCALE, [float(loc), float(scale)])])
