"""Replication programs: what a built-in ``single_simulation`` does, in the kernels' vocabulary.

A program knows (a) which device entry point replays it, (b) how many 64-bit draws one replication
consumes when that is fixed (Pi) or that a planning pass must find out (ziggurat normals).
``run_sequential`` / ``run_blocks`` bind a program to the reference's two stream layouts
(core.py:583-597 and :622-663 of the reference).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _device as D
from . import _native as N
from . import _streams as S


@dataclass
class ReplicationProgram:
    kind: str  # "pi" | "normals"
    # pi
    n_points: int = 0
    antithetic: bool = False
    # normals
    n_steps: int = 0
    specs: list = field(default_factory=list)  # [(mode, params)], common random numbers when > 1

    def execute(self, dl: D.DeviceLayout):
        """-> (float64[n] device tensor of results, int64[n_streams] device tensor of draws per stream | fixed int)"""
        if self.kind == "pi":
            _, out = D.pi_hits(dl, self.n_points, self.antithetic)
            m = self.n_points // 2 if self.antithetic else self.n_points
            per_sim = 2 * m + (2 if (self.antithetic and 2 * m < self.n_points) else 0)
            return out, per_sim
        plan = D.normal_plan(dl, self.n_steps)
        out = D.gbm_terminal(dl, self.n_steps, plan, self.specs)
        return (out[0] if len(self.specs) == 1 else out), plan.stream_end


def pi_program(n_points: int, antithetic: bool) -> ReplicationProgram:
    n_points = int(n_points)
    if n_points <= 0:
        raise ValueError("n_points must be positive")
    return ReplicationProgram("pi", n_points=n_points, antithetic=bool(antithetic))


def normals_program(n_steps: int, specs) -> ReplicationProgram:
    n_steps = int(n_steps)
    if n_steps <= 0:
        raise ValueError("n_steps must be positive")
    return ReplicationProgram("normals", n_steps=n_steps, specs=list(specs))


def run_sequential(prog: ReplicationProgram, rng: np.random.Generator, n: int):
    """One PCG64 stream, replications back to back.  Returns (device results, draws consumed)."""
    dl = D.DeviceLayout.of(S.sequential_layout(rng, n))
    values, used = prog.execute(dl)
    draws = used * n if isinstance(used, int) else int(used.cpu().numpy()[0])
    return values, draws


def run_blocks(prog: ReplicationProgram, blocks, children, n: int, group=None):
    """One Philox stream per spawned block; with ``group`` each rank replays a contiguous range of blocks."""
    lay = S.StreamLayout(S.GEN_PHILOX, S.philox_records(children, blocks), n, blocks[0][1] - blocks[0][0])
    if group is not None:
        import torch.distributed as dist

        lay = lay.shard(dist.get_rank(group), dist.get_world_size(group))
    if lay.n_total == 0:
        import torch

        return torch.empty(0, dtype=torch.float64, device="cuda")
    values, _ = prog.execute(D.DeviceLayout.of(lay))
    return values


__all__ = ["ReplicationProgram", "pi_program", "normals_program", "run_sequential", "run_blocks", "N"]
