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


    def simulate_slices(self, n_paths: int, *, initial_value: float = 10_000.0, annual_return: float = 0.07,
                        volatility: float = 0.20, years: int = 10, slice_stride: int = 21, parallel: bool = False,
                        n_workers: int | None = None, dtype="float64", time_major: bool = True, device: bool = False):
        """EXTENSION (the reference stores no wealth paths, ``sims/portfolio.py:77-84``; the closest reference shape is
        ``BlackScholesPathSimulation.simulate_paths``, ``sims/black_scholes.py:403-418``): GBM wealth of ``n_paths``
        paths at steps ``0, slice_stride, 2*slice_stride, ... <= int(years*252)``.

        The paths are the replications of ``run(n_paths, parallel=parallel, n_workers=n_workers, years=years, ...)``
        with the same seed -- same stream layout, same normals -- so the LAST slice equals that run's results when
        ``slice_stride`` divides the step count.  Like ``run`` it consumes the generator (sequential) or spawns the
        block children (parallel).  Returns ``(n_slices, n_paths)`` when ``time_major`` (the coalesced layout, default)
        else ``(n_paths, n_slices)``; ``dtype`` ``"float64"`` or ``"float32"`` (values computed in fp64, narrowed by the
        store); a host ndarray, or the CUDA tensor with ``device=True``."""
        import multiprocessing as mp

        import torch

        from .. import _device as D
        from .. import _streams as S

        if n_paths <= 0:
            raise ValueError("n_paths must be positive")
        if slice_stride <= 0:
            raise ValueError("slice_stride must be positive")
        tdtype = {"float64": torch.float64, "float32": torch.float32}.get(str(dtype).replace("torch.", ""))
        if tdtype is None:
            raise ValueError("dtype must be 'float64' or 'float32'")
        n_steps = int(years / (1.0 / 252.0))
        if n_steps == 0:
            out = torch.full((1, n_paths) if time_major else (n_paths, 1), float(initial_value), dtype=tdtype, device="cuda")
            return out if device else out.cpu().numpy()
        if n_workers is None:
            n_workers = mp.cpu_count()
        if parallel and n_workers > 1 and n_paths >= self._PARALLEL_THRESHOLD:
            blocks, children = self._prepare_parallel_blocks(n_paths, n_workers)
            lay = S.StreamLayout(S.GEN_PHILOX, S.philox_records(children, blocks), n_paths, blocks[0][1] - blocks[0][0])
            sequential = False
        else:
            lay = S.sequential_layout(self.rng, n_paths)
            sequential = True
        dl = D.DeviceLayout.of(lay)
        plan = D.normal_plan(dl, n_steps)
        out = D.gbm_slices(dl, n_steps, plan, N.PORTFOLIO_GBM, [initial_value, annual_return, volatility],
                           int(slice_stride), time_major=time_major, dtype=tdtype)
        if sequential:
            S.advance_keeping_buffer(self.rng, int(plan.stream_end.cpu().numpy()[0]))
        return out if device else out.cpu().numpy()


PortfolioSimulation._program_hook_owner = PortfolioSimulation
