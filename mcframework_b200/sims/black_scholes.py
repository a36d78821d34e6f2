"""Black-Scholes Monte Carlo: European / path-max "american" payoffs, GBM paths, Longstaff-Schwartz,
bump-and-revalue Greeks -- reference ``mcframework/sims/black_scholes.py``.

Everything numeric runs in the CUDA kernels: payoffs in ``mcf_gbm_terminal`` (one thread per
replication, log-price accumulated in registers), paths in ``mcf_gbm_paths``, the regression in
``mcf_lsm_*``.  ``calculate_greeks`` evaluates all bumped parameter sets on ONE replay of the
normals (the reference re-seeds with 42 before each of its eight runs, :286-353, so they share
their random numbers by construction).
"""
from __future__ import annotations

import multiprocessing as mp
from typing import Optional

import numpy as np
from numpy.random import Generator

from .. import _device as D
from .. import _native as N
from .. import _programs
from .. import _streams as S
from ..core import MonteCarloSimulation
from ._hooks import one_replication

__all__ = ["_european_payoff", "_simulate_gbm_path", "_american_exercise_lsm", "BlackScholesSimulation",
           "BlackScholesPathSimulation"]


def _check_option_type(option_type: str) -> None:
    if option_type not in ("call", "put"):
        raise ValueError(f"option_type must be 'call' or 'put', got '{option_type}'")


def _european_payoff(S_T: float, K: float, option_type: str) -> float:
    """max(S_T - K, 0) for a call, max(K - S_T, 0) for a put (:23-51) -- a scalar, stays on the host."""
    _check_option_type(option_type)
    return max(S_T - K, 0.0) if option_type == "call" else max(K - S_T, 0.0)


def _paths_from(rng: Generator, n_paths: int, S0: float, r: float, sigma: float, T: float, n_steps: int,
                time_major: bool = False):
    """``n_paths`` consecutive GBM paths drawn from ``rng`` (device tensor); ``rng`` is advanced accordingly."""
    kind, rec = S.generator_record(rng, 0, int(n_paths))
    dl = D.DeviceLayout.of(S.StreamLayout(kind, rec, int(n_paths), 0))
    plan = D.normal_plan(dl, int(n_steps))
    paths = D.gbm_paths(dl, int(n_steps), plan, S0, r, sigma, T, time_major=time_major)
    S.consume_draws(rng, int(plan.stream_end.cpu().numpy()[0]))
    return paths


def _simulate_gbm_path(S0: float, r: float, sigma: float, T: float, n_steps: int, rng: Generator) -> np.ndarray:
    """One path ``S_0..S_T`` of ``n_steps + 1`` points from ``rng`` (:54-106)."""
    return _paths_from(rng, 1, S0, r, sigma, T, n_steps)[0].cpu().numpy()


def _american_exercise_lsm(paths, K: float, r: float, dt: float, option_type: str, *, time_major: bool = False) -> float:
    """Longstaff-Schwartz price from a path matrix (:109-193).

    ``paths``: ``(n_paths, n_steps+1)`` ndarray or CUDA tensor (what ``simulate_paths`` returns), or the
    transposed time-major layout with ``time_major=True`` (saves the device transpose)."""
    from .. import _stats_device as SD

    _check_option_type(option_type)
    import torch

    N.lib()
    if isinstance(paths, torch.Tensor):
        dev = paths.to(device="cuda", dtype=torch.float64).contiguous()
    else:
        dev = torch.from_numpy(np.ascontiguousarray(paths, dtype=np.float64)).to("cuda")
    if dev.ndim != 2:
        raise ValueError("paths must be a 2-D array")
    tm = dev if time_major else SD.transpose(dev)
    price = SD.lsm_price(tm, float(K), float(r), float(dt), option_type == "put")
    return float(price.cpu().numpy()[0])


def _bs_mode(option_type: str, exercise_type: str) -> int:
    _check_option_type(option_type)
    if exercise_type not in ("european", "american"):
        raise ValueError(f"exercise_type must be 'european' or 'american', got '{exercise_type}'")
    if exercise_type == "european":
        return N.BS_EURO_CALL if option_type == "call" else N.BS_EURO_PUT
    return N.BS_AMER_CALL if option_type == "call" else N.BS_AMER_PUT


class BlackScholesSimulation(MonteCarloSimulation):
    """Option value per replication (:196-255)."""

    def __init__(self, name: str = "Black-Scholes Option Pricing"):
        super().__init__(name)

    def _replication_program(self, *, S0: float = 100.0, K: float = 100.0, T: float = 1.0, r: float = 0.05,
                             sigma: float = 0.20, option_type: str = "call", exercise_type: str = "european",
                             n_steps: int = 252, **_ignored):
        return _programs.normals_program(n_steps, [(_bs_mode(option_type, exercise_type), [S0, K, T, r, sigma])])

    def single_simulation(self, *, S0: float = 100.0, K: float = 100.0, T: float = 1.0, r: float = 0.05,
                          sigma: float = 0.20, option_type: str = "call", exercise_type: str = "european",
                          n_steps: int = 252, _rng: Optional[Generator] = None, **kwargs) -> float:
        prog = self._replication_program(S0=S0, K=K, T=T, r=r, sigma=sigma, option_type=option_type,
                                         exercise_type=exercise_type, n_steps=n_steps)
        return one_replication(prog, self._rng(_rng, self.rng))

    def calculate_greeks(self, n_simulations: int, S0: float = 100.0, K: float = 100.0, T: float = 1.0, r: float = 0.05,
                         sigma: float = 0.20, option_type: str = "call", exercise_type: str = "european",
                         n_steps: int = 252, parallel: bool = False, bump_pct: float = 0.01,
                         time_bump_days: float = 1.0, *, n_workers: int | None = None) -> dict[str, float]:
        """Central-difference delta, gamma, vega, theta, rho on common random numbers (:257-374)."""
        from .. import _stats_device as SD

        if n_simulations <= 0:
            raise ValueError("n_simulations must be positive")
        mode = _bs_mode(option_type, exercise_type)
        saved = self.rng.bit_generator.state if self.rng else None
        dS, dsig, dT = S0 * bump_pct, sigma * bump_pct, time_bump_days / 365.0
        dr = r * bump_pct if r > 0 else 0.0001
        with_theta = T > dT
        sets = [[S0, K, T, r, sigma], [S0 + dS, K, T, r, sigma], [S0 - dS, K, T, r, sigma], [S0, K, T, r, sigma + dsig],
                [S0, K, T, r, sigma - dsig], [S0, K, T, r + dr, sigma], [S0, K, T, r - dr, sigma]]
        if with_theta:
            sets.append([S0, K, T - dT, r, sigma])

        self.set_seed(42)  # the reference re-seeds with 42 before EVERY revaluation: one stream layout serves all
        if n_workers is None:
            n_workers = mp.cpu_count()  # what run(parallel=True) defaults to in the reference (core.py:547-548)
        group = None
        if parallel and n_workers > 1 and n_simulations >= self._PARALLEL_THRESHOLD:
            blocks, children = self._prepare_parallel_blocks(n_simulations, n_workers)
            lay = S.StreamLayout(S.GEN_PHILOX, S.philox_records(children, blocks), n_simulations, blocks[0][1] - blocks[0][0])
            group = self._shard_group()
            if group is not None:
                import torch.distributed as dist

                lay = lay.shard(dist.get_rank(group), dist.get_world_size(group))
        else:
            lay = S.sequential_layout(self.rng, n_simulations)
        if lay.n_total == 0:  # more ranks than blocks: nothing to replay here, but the moment collectives still run
            import torch

            empty = torch.empty(0, dtype=torch.float64, device="cuda")
            means = [s.mean for s in SD.moments_many([empty] * len(sets), group=group)]
        else:
            dl = D.DeviceLayout.of(lay)

            def replay(plan):
                if exercise_type == "european":  # all parameter sets on ONE replay of the normals
                    values = D.gbm_terminal(dl, int(n_steps), plan, [(mode, p) for p in sets])
                    return [values[k] for k in range(len(sets))]
                # path-dependent payoff: one replay per parameter set, same plan
                return [D.gbm_terminal(dl, int(n_steps), plan, [(mode, p)])[0] for p in sets]

            _, values = D.with_plan(dl, int(n_steps), replay)
            means = [s.mean for s in SD.moments_many(values, group=group)]
        v0, v_up, v_dn, v_sig_up, v_sig_dn, v_r_up, v_r_dn = means[:7]
        greeks = {
            "price": float(v0),
            "delta": float((v_up - v_dn) / (2 * dS)),
            "gamma": float((v_up - 2 * v0 + v_dn) / (dS * dS)),
            "vega": float((v_sig_up - v_sig_dn) / (2 * dsig) * 0.01),
            "theta": float((means[7] - v0) / dT / 365.0) if with_theta else 0.0,
            "rho": float((v_r_up - v_r_dn) / (2 * dr) * 0.01),
        }
        if saved is not None and self.rng is not None:
            self.rng.bit_generator.state = saved
        return greeks


class BlackScholesPathSimulation(MonteCarloSimulation):
    """Terminal value of a stored GBM path, and bulk path generation (:377-418)."""

    def __init__(self, name: str = "Black-Scholes Path Simulation"):
        super().__init__(name)

    def _replication_program(self, *, S0: float = 100.0, r: float = 0.05, sigma: float = 0.20, T: float = 1.0,
                             n_steps: int = 252, **_ignored):
        return _programs.normals_program(n_steps, [(N.PATH_TERMINAL, [S0, r, sigma, T])])

    def single_simulation(self, *, S0: float = 100.0, r: float = 0.05, sigma: float = 0.20, T: float = 1.0,
                          n_steps: int = 252, _rng: Optional[Generator] = None, **kwargs) -> float:
        prog = self._replication_program(S0=S0, r=r, sigma=sigma, T=T, n_steps=n_steps)
        return one_replication(prog, self._rng(_rng, self.rng))

    def simulate_paths(self, n_paths: int, S0: float = 100.0, r: float = 0.05, sigma: float = 0.20, T: float = 1.0,
                       n_steps: int = 252, *, device: bool = False, time_major: bool = False):
        """``(n_paths, n_steps+1)`` matrix of paths drawn back to back from ``self.rng`` (:403-418).

        ``device=True`` keeps the matrix in HBM (optionally time-major, the layout the LSM sweep reads)."""
        paths = _paths_from(self.rng, n_paths, S0, r, sigma, T, n_steps, time_major=time_major)
        return paths if device else paths.cpu().numpy()


BlackScholesSimulation._program_hook_owner = BlackScholesSimulation
BlackScholesPathSimulation._program_hook_owner = BlackScholesPathSimulation
