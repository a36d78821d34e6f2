"""Device reductions behind the StatsEngine metrics and the Longstaff-Schwartz sweep (C ABI wrappers).

Inputs are ``torch`` CUDA ``float64`` tensors (ownership only); every number is produced by the
kernels in ``csrc/mcf_stats.cu``.  Host code here only (a) sizes workspaces, (b) turns NumPy's
quantile definition into order-statistic ranks and applies its two-point ``_lerp`` to the two
selected values, (c) merges per-rank partials for the multi-GPU path.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _native as N
from ._device import LAUNCHES, _ptr, _stream_ptr, _torch
from ._streams import consume_draws, generator_record

MOMENTS_OUT = 12


# --------------------------------------------------------------------------------------- moments
@dataclass
class MomentSummary:
    """One-pass summary of a sample (finite values only; non-finite values are counted apart)."""

    n: float
    n_nan: float
    n_posinf: float
    n_neginf: float
    mean: float
    m2: float
    m3: float
    m4: float
    sq_target: float
    total: float
    reserved0: float = 0.0
    reserved1: float = 0.0

    @classmethod
    def from_array(cls, a) -> "MomentSummary":
        return cls(*[float(v) for v in a])

    def to_array(self) -> np.ndarray:
        return np.array([self.n, self.n_nan, self.n_posinf, self.n_neginf, self.mean, self.m2, self.m3, self.m4,
                         self.sq_target, self.total, 0.0, 0.0], dtype=np.float64)

    @property
    def n_nonfinite(self) -> float:
        return self.n_nan + self.n_posinf + self.n_neginf


def merge_summaries(parts) -> MomentSummary:
    """Pairwise (Chan / Pebay) combination of per-rank summaries, in rank order."""
    parts = list(parts)
    acc = parts[0]
    for b in parts[1:]:
        a = acc
        if b.n == 0:
            mean, m2, m3, m4, n = a.mean, a.m2, a.m3, a.m4, a.n
        elif a.n == 0:
            mean, m2, m3, m4, n = b.mean, b.m2, b.m3, b.m4, b.n
        else:
            n = a.n + b.n
            d = b.mean - a.mean
            dn = d / n
            dn2 = dn * dn
            mean = a.mean + dn * b.n
            m2 = a.m2 + b.m2 + d * dn * a.n * b.n
            m3 = a.m3 + b.m3 + d * dn2 * a.n * b.n * (a.n - b.n) + 3.0 * dn * (a.n * b.m2 - b.n * a.m2)
            m4 = (a.m4 + b.m4 + d * dn * dn2 * a.n * b.n * (a.n * a.n - a.n * b.n + b.n * b.n)
                  + 6.0 * dn2 * (a.n * a.n * b.m2 + b.n * b.n * a.m2) + 4.0 * dn * (a.n * b.m3 - b.n * a.m3))
        acc = MomentSummary(n, a.n_nan + b.n_nan, a.n_posinf + b.n_posinf, a.n_neginf + b.n_neginf, mean, m2, m3, m4,
                            a.sq_target + b.sq_target, a.total + b.total)
    return acc


def moments_raw(x, target: float = 0.0):
    """Launch the one-pass moments kernel; returns the 12-double device tensor (no sync)."""
    torch = _torch()
    lib = N.lib()
    n = int(x.numel())
    ws = torch.empty(int(lib.mcf_moments_workspace_bytes(n)), dtype=torch.uint8, device=x.device)
    out = torch.empty(MOMENTS_OUT, dtype=torch.float64, device=x.device)
    rc = lib.mcf_moments(_ptr(x), n, float(target), _ptr(out), _ptr(ws), int(ws.numel()), _stream_ptr())
    N.check("mcf_moments", rc)
    LAUNCHES["stats"] += 2
    return out


def moments(x, target: float = 0.0, group=None) -> MomentSummary:
    """Summary of the sample held in ``x``; with ``group`` (a torch.distributed process group) the
    sample is the concatenation of every rank's ``x`` (one small all-gather of 12 doubles per rank)."""
    torch = _torch()
    if x.numel() == 0:
        out = torch.zeros(MOMENTS_OUT, dtype=torch.float64, device=x.device)
    else:
        out = moments_raw(x, target)
    if group is None:
        return MomentSummary.from_array(out.cpu().numpy())
    import torch.distributed as dist

    world = dist.get_world_size(group)
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out, group=group)
    return merge_summaries(MomentSummary.from_array(g.cpu().numpy()) for g in gathered)


def moments_many(xs, target: float = 0.0, group=None):
    """``[moments(x, target, group) for x in xs]`` with ONE device->host copy (and one all-gather) for all of them:
    the eight revaluations of ``calculate_greeks`` are read back together."""
    torch = _torch()
    xs = list(xs)
    if not xs:
        return []
    outs = [torch.zeros(MOMENTS_OUT, dtype=torch.float64, device=x.device) if x.numel() == 0 else moments_raw(x, target)
            for x in xs]
    both = torch.stack(outs)
    if group is None:
        host = both.cpu().numpy()
        return [MomentSummary.from_array(row) for row in host]
    import torch.distributed as dist

    world = dist.get_world_size(group)
    gathered = [torch.empty_like(both) for _ in range(world)]
    dist.all_gather(gathered, both, group=group)
    host = torch.stack(gathered).cpu().numpy()  # (world, len(xs), 12)
    return [merge_summaries(MomentSummary.from_array(host[r, k]) for r in range(world)) for k in range(len(xs))]


# ---------------------------------------------------------------------------------------- select
#: samples of at least this many values compact the surviving candidates after the second radix pass
SELECT_COMPACT_MIN_N = 1 << 20


def select(x, ranks, omit_nonfinite: bool = False, group=None):
    """Order statistics ``sorted(x)[ranks]`` (float64 device tensor) by radix selection.

    With ``group`` every pass's histogram is all-reduced, so ``ranks`` index the union of all ranks' data."""
    torch = _torch()
    lib = N.lib()
    ranks = [int(r) for r in ranks]
    k = len(ranks)
    nbytes = int(lib.mcf_select_workspace_bytes())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty(k, dtype=torch.float64, device=x.device)
    arr = (C.c_int64 * k)(*ranks)
    n = int(x.numel())
    if group is None:
        if n >= SELECT_COMPACT_MIN_N:  # large samples: passes 3..8 read only the candidates left after two digits
            cbytes = int(lib.mcf_select_compact_bytes(n))
            scratch = torch.empty(cbytes, dtype=torch.uint8, device=x.device)
            rc = lib.mcf_select_compacting(_ptr(x), n, int(bool(omit_nonfinite)), arr, k, _ptr(out), _ptr(ws), nbytes,
                                           _ptr(scratch), cbytes, _stream_ptr())
            N.check("mcf_select_compacting", rc)
            LAUNCHES["stats"] += 4 + 2 * int(lib.mcf_select_passes())
            del scratch
            return out
        rc = lib.mcf_select(_ptr(x), n, int(bool(omit_nonfinite)), arr, k, _ptr(out), _ptr(ws), nbytes, _stream_ptr())
        N.check("mcf_select", rc)
        LAUNCHES["stats"] += 2 + 2 * int(lib.mcf_select_passes())
        return out
    import torch.distributed as dist

    off, hb = int(lib.mcf_select_hist_offset()), int(lib.mcf_select_hist_bytes())
    hist = ws[off:off + hb].view(torch.int64)
    N.check("mcf_select_begin", lib.mcf_select_begin(arr, k, _ptr(ws), nbytes, _stream_ptr()))
    for _ in range(int(lib.mcf_select_passes())):
        if n > 0:
            N.check("mcf_select_hist", lib.mcf_select_hist(_ptr(x), n, int(bool(omit_nonfinite)), _ptr(ws), _stream_ptr()))
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        N.check("mcf_select_choose", lib.mcf_select_choose(_ptr(ws), _stream_ptr()))
    N.check("mcf_select_finish", lib.mcf_select_finish(_ptr(ws), _ptr(out), _stream_ptr()))
    LAUNCHES["stats"] += 2 + 2 * int(lib.mcf_select_passes())
    return out


#: samples of at least this many values try the one-pass bracketed selection first (exact; falls back when undecided)
SELECT_BRACKET_MIN_N = 1 << 23


def select_bracketed(x, ranks):
    """``sorted(x)[ranks]`` for a LARGE NaN-free sample in one pass over ``x`` (``mcf_select_bracketed``): float64 device
    tensor in the order of ``ranks``, or ``None`` when the one-pass form could not decide (the caller then runs the radix
    selection).  At most 16 distinct ranks."""
    torch = _torch()
    lib = N.lib()
    n = int(x.numel())
    uniq = sorted({int(r) for r in ranks})
    if n < SELECT_BRACKET_MIN_N or not uniq or len(uniq) > 16 or uniq[0] < 0 or uniq[-1] >= n:
        return None
    k = len(uniq)
    nbytes = int(lib.mcf_select_workspace_bytes())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    sbytes = int(lib.mcf_select_bracket_bytes(n))
    scratch = torch.empty(sbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty(k + 1, dtype=torch.float64, device=x.device)  # [k] carries the "undecided" flag back in the same copy
    flag = out[k:].view(torch.int32)
    arr = (C.c_int64 * k)(*uniq)
    rc = lib.mcf_select_bracketed(_ptr(x), n, arr, k, _ptr(out), _ptr(ws), nbytes, _ptr(scratch), sbytes, _ptr(flag), _stream_ptr())
    N.check("mcf_select_bracketed", rc)
    LAUNCHES["stats"] += 24 + 2 * int(lib.mcf_select_passes())
    host = out.cpu()
    if int(host[k:].view(torch.int32)[0]) != 0:
        return None
    pos = {r: i for i, r in enumerate(uniq)}
    return host[:k][[pos[int(r)] for r in ranks]]


def quantile_plan(n: int, percentiles):
    """NumPy's default ('linear') quantile definition as order-statistic ranks + interpolation weights
    (numpy/lib/_function_base_impl.py: 'linear' virtual index ``(n-1)*q``, ``_get_indexes``, ``_get_gamma``)."""
    q = np.true_divide(np.asarray(list(percentiles), dtype=np.float64), 100)
    virt = (n - 1) * q
    prev = np.floor(virt)
    nxt = prev + 1
    above = virt >= n - 1
    prev[above] = n - 1
    nxt[above] = n - 1
    below = virt < 0
    prev[below] = 0
    nxt[below] = 0
    gamma = virt - np.floor(virt)
    return prev.astype(np.int64), nxt.astype(np.int64), gamma


def lerp_like_numpy(a, b, t):
    """numpy/lib/_function_base_impl.py ``_lerp`` on float64 vectors."""
    a, b, t = (np.asarray(v, dtype=np.float64) for v in (a, b, t))
    diff = b - a
    out = a + diff * t
    hi = t >= 0.5
    out[hi] = (b - diff * (1 - t))[hi]
    return out


def percentiles(x, pcts, n_valid: int | None = None, has_nan: bool = False, omit_nonfinite: bool = False, group=None):
    """``np.percentile(x, pcts)`` (method 'linear') with the sort replaced by device radix selection.

    ``n_valid``: number of values taking part (all of them, or the finite ones with ``omit_nonfinite``).
    ``has_nan`` (and not omitting): NumPy returns NaN for every percentile."""
    pcts = list(pcts)
    q = np.asarray(pcts, dtype=np.float64)
    if q.size and not (np.all(q >= 0.0) and np.all(q <= 100.0)):  # NaN fails both tests, as in np.percentile
        raise ValueError("Percentiles must be in the range [0, 100]")
    n = int(x.numel()) if n_valid is None else int(n_valid)
    if n == 0 or (has_nan and not omit_nonfinite):
        return np.full(len(pcts), np.nan)
    prev, nxt, gamma = quantile_plan(n, pcts)
    ranks = np.concatenate([prev, nxt])
    if group is None and not omit_nonfinite and n == int(x.numel()):  # large NaN-free samples: one pass over x
        got = select_bracketed(x, ranks)
        if got is not None:
            vals = got.numpy()
            return lerp_like_numpy(vals[:len(pcts)], vals[len(pcts):], gamma)
    out = np.empty(len(ranks), dtype=np.float64)
    for i in range(0, len(ranks), 32):  # MCF_SELECT_MAX_RANKS per launch
        out[i:i + 32] = select(x, ranks[i:i + 32], omit_nonfinite, group).cpu().numpy()
    k = len(pcts)
    return lerp_like_numpy(out[:k], out[k:], gamma)


# ------------------------------------------------------------------------------------- bootstrap
def _advance_generator(gen: np.random.Generator, slots_consumed: int) -> None:
    """Leave ``gen`` in the state NumPy's own ``integers`` call would have left it."""
    bg = gen.bit_generator
    st = bg.state
    has = int(st["has_uint32"])
    t = slots_consumed - has  # slots taken from fresh 64-bit draws
    if t <= 0:
        if slots_consumed == 1 and has:
            st["has_uint32"] = 0
            bg.state = st
        return
    last = consume_draws(gen, (t + 1) // 2)  # the draw whose halves were (partly) used
    st = bg.state
    st["has_uint32"] = 1 if (t & 1) else 0
    st["uinteger"] = last >> 32  # NumPy keeps the (possibly already consumed) upper half in the state
    bg.state = st


def rejection_probability(n: int) -> float:
    return float(((1 << 32) % n) / float(1 << 32))


#: populations above this many elements use the binned (L2-resident regions) gather path
BINNED_MIN_N = 12_000_000
#: how many binned bootstraps were decided by the single-pass form / recomputed or run by the two-pass form (tests, bench)
BOOT_FORMS = {"single_pass": 0, "two_pass": 0}


def _slot_window(need: int, p: float, extra_sigma: float) -> int:
    """Slots of the single index stream that hold ``need`` accepted ones, + ``extra_sigma`` standard deviations."""
    if p <= 0:
        return need
    expect = need / (1.0 - p)
    return int(math.ceil(expect + extra_sigma * math.sqrt(expect * p / max(1e-300, (1.0 - p))) + 64))


def _default_resamples_per_batch(lib, x, n: int, B: int, region_shift: int) -> int:
    """Resamples binned per batch: 64 when the two queue sets (8 bytes per index and batch: 53 GB at n = 1e8) take no more
    than a third of the free device memory, else 32, else 16.  Measured at n = 1e8, B = 512: 16 -> 1452, 32 -> 1591,
    48 -> 1646, 64 -> 1672 resamples/s (fewer, larger launches; the queues are scratch of the call)."""
    torch = _torch()
    free = torch.cuda.mem_get_info(x.device)[0] if x.is_cuda else 0
    for cand in (64, 32):
        rpb = max(1, min(cand, B))  # (a batch never needs more queues than there are resamples)
        if int(lib.mcf_bootstrap_binned_workspace_bytes(n, rpb, int(region_shift))) * 3 <= free:
            return rpb
    return max(1, min(16, B))


def _bootstrap_single_pass(x, n, B, kind, d_rec, p, rank, world, group, chunk_slots, region_shift, resamples_per_batch,
                           sums, end_slot) -> bool:
    """Binned bootstrap without the count pass (``mcf_bootstrap_fused_phase1/2``): CTAs whose estimated ordinal range lies
    inside one resample bin and count in one pass, the rest get exact ordinals.  False = undecided, nothing usable in
    ``sums`` (the caller recomputes with the two-pass form)."""
    torch = _torch()
    lib = N.lib()
    total_slots = _slot_window(B * n, p, 8.0)
    lo, hi = total_slots * rank // world, total_slots * (rank + 1) // world
    nslots = hi - lo
    cs = chunk_slots or int(min(8192, max(64, (nslots // (148 * 2048 * 2)) // 8 * 8)))
    if 256 * cs > n:
        return False
    nbytes = int(lib.mcf_bootstrap_workspace_bytes(nslots, cs))
    bbytes = int(lib.mcf_bootstrap_binned_workspace_bytes(n, int(resamples_per_batch), int(region_shift)))
    if nbytes < 0 or bbytes < 0:
        return False
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    bws = torch.empty(bbytes, dtype=torch.uint8, device=x.device)
    total = torch.zeros(1, dtype=torch.int64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    sp = _stream_ptr()
    rc = lib.mcf_bootstrap_fused_phase1(kind, _ptr(d_rec), _ptr(x), n, B, lo, hi, total_slots, cs, _ptr(ws), nbytes,
                                        int(resamples_per_batch), int(region_shift), _ptr(bws), bbytes, _ptr(sums), _ptr(total), sp)
    ok = rc == N.OK
    if group is not None:
        import torch.distributed as dist

        totals = [torch.empty_like(total) for _ in range(world)]
        dist.all_gather(totals, total, group=group)
        base = sum(int(t.item()) for t in totals[:rank])
    else:
        base = 0
    if ok:
        rc = lib.mcf_bootstrap_fused_phase2(kind, _ptr(d_rec), _ptr(x), n, B, lo, hi, total_slots, cs, base, _ptr(ws), nbytes,
                                            int(resamples_per_batch), int(region_shift), _ptr(bws), bbytes, _ptr(sums),
                                            _ptr(end_slot), _ptr(flag), sp)
        ok = rc == N.OK
    LAUNCHES["stats"] += 6
    bad = torch.tensor([0 if ok else 1], dtype=torch.int32, device=x.device) + flag
    if group is not None:
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(end_slot, op=dist.ReduceOp.MAX, group=group)
    if int(bad.item()) != 0 or int(end_slot.item()) == 0:
        return False
    if group is not None:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return True


def bootstrap_means(x, n_resamples: int, gen: np.random.Generator, group=None, advance: bool = True,
                    chunk_slots: int | None = None, binned: bool | None = None, region_shift: int = 0,
                    resamples_per_batch: int | None = None, single_pass: bool = True):
    """``x[rng.integers(0, n, size=(B, n))].mean(axis=1)`` (stats_engine.py:1039-1041) without the index
    matrix: float64[B] device tensor.  ``gen`` is advanced exactly as NumPy would have advanced it.

    With ``group`` the slot range of the stream is sharded across ranks (every rank holds all of ``x``);
    two small collectives combine the result (all-gather of accepted counts, all-reduce of B sums).

    Populations larger than L2 (``binned``; default from ``BINNED_MIN_N``) first queue the indices per (resample,
    64 MB region of ``x``) and then sum region by region: a random gather from an L2-resident region is 3-4x faster
    than one from HBM, where every 8-byte gather moves a 128-byte line."""
    torch = _torch()
    lib = N.lib()
    n, B = int(x.numel()), int(n_resamples)
    if n == 1:  # NumPy draws nothing for a single-element population
        return x.to(torch.float64).expand(B).clone()
    if n < 2 or n >= (1 << 32):
        raise ValueError("bootstrap_means needs 1 <= n < 2**32")
    kind, rec = generator_record(gen)
    d_rec = torch.from_numpy(rec.view(np.uint8).copy()).to(x.device)
    has = int(rec["has_uint32"][0])
    need = B * n
    p = rejection_probability(n)
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
    sums = torch.zeros(B, dtype=torch.float64, device=x.device)
    end_slot = torch.zeros(1, dtype=torch.int64, device=x.device)
    if binned is None:
        binned = n >= BINNED_MIN_N
    if resamples_per_batch is None:
        resamples_per_batch = _default_resamples_per_batch(lib, x, n, B, region_shift) if binned else 32
    if binned and single_pass and p > 0:
        done = _bootstrap_single_pass(x, n, B, kind, d_rec, p, rank, world, group, chunk_slots, region_shift,
                                      resamples_per_batch, sums, end_slot)
        if done:
            BOOT_FORMS["single_pass"] += 1
            N.check("mcf_bootstrap_finalize", lib.mcf_bootstrap_finalize(_ptr(sums), B, n, _stream_ptr()))
            if advance:
                _advance_generator(gen, int(end_slot.item()))
            return sums
        sums.zero_()  # undecided (an 8-sigma miss, a window too small): the two-pass form below recomputes everything
        end_slot.zero_()
    extra_sigma = 8.0
    for _attempt in range(4):
        expect = need / (1.0 - p)
        total_slots = int(math.ceil(expect + extra_sigma * math.sqrt(expect * p / max(1e-300, (1.0 - p))) + 64)) if p > 0 else need
        lo = total_slots * rank // world
        hi = total_slots * (rank + 1) // world
        nslots = hi - lo
        cs = chunk_slots or int(min(8192, max(64, (nslots // (148 * 2048 * 2)) // 8 * 8)))
        nbytes = int(lib.mcf_bootstrap_workspace_bytes(nslots, cs))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        total = torch.zeros(1, dtype=torch.int64, device=x.device)
        rc = lib.mcf_bootstrap_count(kind, _ptr(d_rec), n, lo, hi, cs, _ptr(ws), nbytes, _ptr(total), _stream_ptr())
        N.check("mcf_bootstrap_count", rc)
        LAUNCHES["stats"] += 2
        if group is not None:
            totals = [torch.empty_like(total) for _ in range(world)]
            dist.all_gather(totals, total, group=group)
            totals = [int(t.item()) for t in totals]
        else:
            totals = [int(total.item())]
        if sum(totals) >= need:
            break
        extra_sigma *= 4.0
    else:
        raise N.NativeError("mcf_bootstrap_count", N.ERR_WINDOW_TOO_SMALL, "slot window too small after retries")
    base = sum(totals[:rank])
    if binned and 256 * cs > n:
        binned = False  # a CTA must not span more than two resamples
    if binned:
        BOOT_FORMS["two_pass"] += 1
        nb = int(lib.mcf_bootstrap_blocks(nslots, cs))
        h_bb = np.empty(nb, dtype=np.uint64)
        N.check("mcf_bootstrap_block_base", lib.mcf_bootstrap_block_base(_ptr(ws), nslots, cs, h_bb.ctypes.data_as(C.c_void_p), _stream_ptr()))
        bbytes = int(lib.mcf_bootstrap_binned_workspace_bytes(n, int(resamples_per_batch), int(region_shift)))
        if bbytes < 0:
            N.check("mcf_bootstrap_binned_workspace_bytes", bbytes)
        bws = torch.empty(bbytes, dtype=torch.uint8, device=x.device)
        rc = lib.mcf_bootstrap_accumulate_binned(kind, _ptr(d_rec), _ptr(x), n, B, lo, hi, cs, base, _ptr(ws), nbytes,
                                                 h_bb.ctypes.data_as(C.c_void_p), int(resamples_per_batch), int(region_shift),
                                                 _ptr(bws), bbytes, _ptr(sums), _ptr(end_slot), _stream_ptr())
        N.check("mcf_bootstrap_accumulate_binned", rc)
        LAUNCHES["stats"] += 2
        del bws
    else:
        rc = lib.mcf_bootstrap_accumulate(kind, _ptr(d_rec), _ptr(x), n, B, lo, hi, cs, base, _ptr(ws), nbytes, _ptr(sums),
                                          _ptr(end_slot), _stream_ptr())
        N.check("mcf_bootstrap_accumulate", rc)
        LAUNCHES["stats"] += 2
    if group is not None:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(end_slot, op=dist.ReduceOp.MAX, group=group)
    N.check("mcf_bootstrap_finalize", lib.mcf_bootstrap_finalize(_ptr(sums), B, n, _stream_ptr()))
    if advance:
        _advance_generator(gen, int(end_slot.item()))
    del ws
    _ = has
    return sums


# ------------------------------------------------------------------------------------------- lsm
def transpose(a):
    """(rows, cols) row-major float64 -> (cols, rows) row-major, on device."""
    torch = _torch()
    rows, cols = int(a.shape[0]), int(a.shape[1])
    out = torch.empty((cols, rows), dtype=torch.float64, device=a.device)
    N.check("mcf_transpose", N.lib().mcf_transpose(_ptr(a), rows, cols, _ptr(out), _stream_ptr()))
    LAUNCHES["lsm"] += 1
    return out


class LsmPeerExchange:
    """Peer-mapped exchange buffers of one process group (the GPUs of one box, at most 8): every rank allocates a 2 KB
    buffer, the CUDA IPC handles travel once through ``all_gather_object``, every rank maps every buffer.  The fused
    sharded sweep (``mcf_lsm_sharded``) then exchanges the 8 sums of a step inside its kernel over NVLink."""

    _cache: dict = {}

    def __init__(self, group):
        import socket

        import torch.distributed as dist

        lib = N.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.seq = 0
        self.ok = False
        local, handle = C.c_void_p(), (C.c_ubyte * 64)()
        mine = None
        if self.world <= 8:
            rc = lib.mcf_peer_buffer_create(int(lib.mcf_lsm_exchange_bytes()), C.byref(local), handle)
            mine = (socket.gethostname(), bytes(handle)) if rc == N.OK else None
        infos = [None] * self.world
        dist.all_gather_object(infos, mine, group=group)
        good = all(i is not None for i in infos) and len({i[0] for i in infos}) == 1
        ptrs = []
        if good:
            for r, (_, h) in enumerate(infos):
                if r == self.rank:
                    ptrs.append(local.value)
                    continue
                q = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                if lib.mcf_peer_buffer_open(buf, C.byref(q)) != N.OK:
                    good = False
                    break
                ptrs.append(q.value)
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(good), group=group)  # every rank takes the same path
        self.ok = all(flags)
        self.ptrs = (C.c_void_p * 8)(*(ptrs + [None] * (8 - len(ptrs)))) if self.ok else None

    @classmethod
    def get(cls, group):
        key = id(group)
        if key not in cls._cache:
            cls._cache[key] = cls(group)
        return cls._cache[key]


def lsm_price(paths_tm, K: float, r: float, dt: float, is_put: bool, group=None, return_exercise: bool = False,
              fused: bool = True):
    """Longstaff-Schwartz price from TIME-MAJOR paths ``(n_steps+1, n_paths)`` on device
    (sims/black_scholes.py:109-193).  With ``group`` the paths are sharded by rank; the 8 regression sums of
    every step and the final cash-flow sum are all-reduced."""
    torch = _torch()
    lib = N.lib()
    n_steps, n_paths = int(paths_tm.shape[0]) - 1, int(paths_tm.shape[1])
    nbytes = int(lib.mcf_lsm_workspace_bytes(n_paths))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=paths_tm.device)
    price = torch.empty(1, dtype=torch.float64, device=paths_tm.device)
    ex = torch.empty(n_paths, dtype=torch.int32, device=paths_tm.device) if return_exercise else None
    sp = _stream_ptr()
    if group is None:
        rc = lib.mcf_lsm(_ptr(paths_tm), n_paths, n_steps, float(K), float(r), float(dt), int(bool(is_put)), _ptr(ws),
                         nbytes, _ptr(price), _ptr(ex) if ex is not None else None, sp)
        N.check("mcf_lsm", rc)
        LAUNCHES["lsm"] += n_steps  # fused apply+reduce step kernel, one launch per step (the last CTA solves)
        return (price, ex) if return_exercise else price
    import torch.distributed as dist

    px = LsmPeerExchange.get(group) if fused else None
    if px is not None and px.ok:
        # one fused launch per step and rank; the ranks' sums meet inside the kernel (peer memory over NVLink)
        counts = torch.tensor([float(n_paths)], dtype=torch.float64, device=paths_tm.device)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)  # also orders this sweep behind the previous one
        n_global = int(round(float(counts.item())))
        base = px.seq
        px.seq += n_steps + 2
        rc = lib.mcf_lsm_sharded(_ptr(paths_tm), n_paths, n_global, n_steps, float(K), float(r), float(dt), int(bool(is_put)),
                                 _ptr(ws), nbytes, px.ptrs, px.rank, px.world, base, _ptr(price),
                                 _ptr(ex) if ex is not None else None, sp)
        N.check("mcf_lsm_sharded", rc)
        LAUNCHES["lsm"] += n_steps
        return (price, ex) if return_exercise else price
    sums = ws[:64].view(torch.float64)
    N.check("mcf_lsm_begin", lib.mcf_lsm_begin(_ptr(paths_tm), n_paths, n_steps, float(K), float(r), float(dt), int(bool(is_put)), _ptr(ws), nbytes, sp))
    for t in range(n_steps - 1, 0, -1):
        N.check("mcf_lsm_reduce", lib.mcf_lsm_reduce(_ptr(paths_tm), n_paths, t, float(K), float(r), float(dt), int(bool(is_put)), _ptr(ws), sp))
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        N.check("mcf_lsm_apply", lib.mcf_lsm_apply(_ptr(paths_tm), n_paths, t, float(K), float(r), float(dt), int(bool(is_put)), _ptr(ws), sp))
    psum = torch.empty(1, dtype=torch.float64, device=paths_tm.device)
    N.check("mcf_lsm_finish", lib.mcf_lsm_finish(n_paths, float(r), float(dt), _ptr(ws), _ptr(psum), None,
                                                 _ptr(ex) if ex is not None else None, sp))
    cnt = torch.tensor([float(n_paths)], dtype=torch.float64, device=paths_tm.device)
    both = torch.cat([psum, cnt])
    dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
    price = (both[0] / both[1]).reshape(1)
    LAUNCHES["lsm"] += 1 + 4 * max(0, n_steps - 1) + 3
    return (price, ex) if return_exercise else price


# --------------------------------------------------------------------------------------- helpers
def count_less(x, value: float) -> int:
    """``int(np.sum(x < value))`` on device (one 8-byte read back)."""
    torch = _torch()
    if x.numel() == 0:
        return 0
    cnt = torch.empty(1, dtype=torch.int64, device=x.device)
    N.check("mcf_count_less", N.lib().mcf_count_less(_ptr(x), int(x.numel()), float(value), _ptr(cnt), _stream_ptr()))
    LAUNCHES["stats"] += 1
    return int(cnt.item())


def compact_finite(x):
    """``x[np.isfinite(x)]`` on device, order preserved (nan_policy="omit")."""
    torch = _torch()
    lib = N.lib()
    n = int(x.numel())
    if n == 0:
        return x
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    cnt = torch.empty(1, dtype=torch.int64, device=x.device)
    nbytes = int(lib.mcf_compact_workspace_bytes(n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    N.check("mcf_compact_finite", lib.mcf_compact_finite(_ptr(x), n, _ptr(out), _ptr(cnt), _ptr(ws), nbytes, _stream_ptr()))
    LAUNCHES["stats"] += 3
    return out[: int(cnt.item())]
