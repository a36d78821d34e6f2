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


class HostLoopRequired(RuntimeError):
    """The declared program met a case only the reference's own loop reproduces (a rejected Lemire slot)."""


@dataclass
class ReplicationProgram:
    kind: str  # "pi" | "normals" | "integers" | "constant"
    # pi
    n_points: int = 0
    antithetic: bool = False
    # normals
    n_steps: int = 0
    specs: list = field(default_factory=list)  # [(mode, params)], common random numbers when > 1
    # integers
    low: int = 0
    high: int = 0
    size: int = 0
    op: int = 0      # 0: sum of the values, 1: longest run of `match`
    match: int = 0
    # constant: a replication that draws nothing (a zero-step horizon)
    value: float = 0.0

    def execute(self, dl: D.DeviceLayout):
        """-> (float64[n] device tensor of results, int64[n_streams] device tensor of draws per stream | fixed int)"""
        if self.kind == "constant":
            import torch

            return torch.full((dl.layout.n_total,), self.value, dtype=torch.float64, device="cuda"), 0
        if self.kind == "pi":
            _, out = D.pi_hits(dl, self.n_points, self.antithetic)
            m = self.n_points // 2 if self.antithetic else self.n_points
            per_sim = 2 * m + (2 if (self.antithetic and 2 * m < self.n_points) else 0)
            return out, per_sim
        if self.kind == "integers":
            out, rejected = D.draw_integers(dl, self.low, self.high, self.size, self.op, self.match)
            if int(rejected.item()):
                raise HostLoopRequired("a Lemire slot was rejected: later slots shift, run the hook's own loop")
            return out, SlotCount(self.size)
        plan, out = D.with_plan(dl, self.n_steps, lambda plan: D.gbm_terminal(dl, self.n_steps, plan, self.specs))
        return (out[0] if len(self.specs) == 1 else out), plan.stream_end


class SlotCount(int):
    """32-bit slots (not 64-bit draws) one replication consumes: the generator keeps half a draw buffered."""


def integers_program(low: int, high: int, size: int, op: int = 0, match: int = 0) -> ReplicationProgram:
    low, high, size = int(low), int(high), int(size)
    if size <= 0 or not (2 <= high - low < (1 << 32)):
        raise ValueError("integers program needs size >= 1 and 2 <= high - low < 2**32")
    return ReplicationProgram("integers", low=low, high=high, size=size, op=int(op), match=int(match))


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


def constant_program(value: float) -> ReplicationProgram:
    """Every replication returns ``value`` and leaves its stream untouched (``rng.normal(size=0)`` in the reference)."""
    return ReplicationProgram("constant", value=float(value))


def run_sequential(prog: ReplicationProgram, rng: np.random.Generator, n: int):
    """One PCG64 stream, replications back to back.  Returns (device results, draws consumed)."""
    dl = D.DeviceLayout.of(S.sequential_layout(rng, n))
    values, used = prog.execute(dl)
    if isinstance(used, SlotCount):  # buffered 32-bit slots: leave the generator (and its half-draw buffer) as NumPy would
        from ._stats_device import _advance_generator

        _advance_generator(rng, int(used) * n)
        return values, 0
    draws = used * n if isinstance(used, int) else int(used.cpu().numpy()[0])
    return values, draws


def run_blocks(prog: ReplicationProgram, blocks, children, n: int, group=None):
    """One Philox stream per spawned block; with ``group`` each rank replays a contiguous range of blocks."""
    lay = S.StreamLayout(S.GEN_PHILOX, S.philox_records(children, blocks), n, blocks[0][1] - blocks[0][0])
    if group is not None:
        import torch.distributed as dist

        lay = lay.shard(dist.get_rank(group), dist.get_world_size(group))
    if lay.n_total == 0:  # more ranks than blocks: an empty shard still takes part in the collectives below
        import torch

        values, failed = torch.empty(0, dtype=torch.float64, device="cuda"), False
    else:
        try:
            values, _ = prog.execute(D.DeviceLayout.of(lay))
            failed = False
        except HostLoopRequired:
            values, failed = None, True
    if group is not None:  # every rank takes the same path
        import torch
        import torch.distributed as dist

        flag = torch.tensor([1 if failed else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        failed = bool(int(flag.item()))
    if failed:
        raise HostLoopRequired("a rank met a rejected Lemire slot")
    return values


def run_blocks_pipelined(prog: ReplicationProgram, blocks, children, n: int, n_groups: int = 4, shard=None, host_out=None):
    """``run_blocks`` for a host-resident result: the spawned streams are replayed in ``n_groups`` contiguous groups and
    the device->host copy of group g (pinned memory, side stream) overlaps the kernels of group g+1, so only the last
    group's copy is exposed; the plan workspace shrinks by the same factor.  Streams are independent, so the values are
    those of ``run_blocks`` bit for bit.  ``shard = (rank, world)``: only this rank's contiguous range of blocks is
    replayed; ``host_out``: a page-locked float64 host tensor for exactly that range (a slice of the shared result
    buffer of a multi-GPU run) instead of a fresh pinned allocation.  Returns (device results, host ndarray)."""
    import torch

    lay = S.StreamLayout(S.GEN_PHILOX, S.philox_records(children, blocks), n, blocks[0][1] - blocks[0][0])
    if shard is not None:
        lay = lay.shard(*shard)
        n = lay.n_total
    values = torch.empty(n, dtype=torch.float64, device="cuda")
    host = host_out if host_out is not None else torch.empty(n, dtype=torch.float64, pin_memory=True)
    assert host.numel() == n, (host.numel(), n)
    if n == 0:
        return values, host.numpy()
    n_groups = max(1, min(int(n_groups), lay.n_streams))
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    at = 0
    for g in range(n_groups):
        part = lay.shard(g, n_groups)
        if part.n_total == 0:
            continue
        out, _ = prog.execute(D.DeviceLayout.of(part))
        piece = values[at:at + part.n_total]
        piece.copy_(out)
        del out
        done = torch.cuda.Event()
        done.record(main)
        side.wait_event(done)
        with torch.cuda.stream(side):
            host[at:at + part.n_total].copy_(piece, non_blocking=True)
        at += part.n_total
    assert at == n, (at, n)
    side.synchronize()
    return values, host.numpy()


def shard_sizes(blocks, world: int):
    """Replications per rank of the contiguous block sharding (what ``StreamLayout.shard`` gives every rank)."""
    k = len(blocks)
    sizes = []
    for r in range(world):
        lo, hi = r * k // world, (r + 1) * k // world
        sizes.append(sum(b - a for a, b in blocks[lo:hi]))
    return sizes


__all__ = ["ReplicationProgram", "HostLoopRequired", "SlotCount", "pi_program", "normals_program", "integers_program",
           "constant_program", "run_sequential", "run_blocks", "run_blocks_pipelined", "shard_sizes", "N"]
