"""TEST INFRASTRUCTURE ONLY -- an oracle-backed stand-in for the device entry points.

Purpose: run the REFERENCE's own test-suite (``/root/reference/tests``, read in place, never copied) against
``mcframework_b200``'s host logic on a box without a GPU, as a drop-in check of everything above the C ABI:
signatures, validation, error taxonomy, stream bookkeeping, statistics assembly.  ``install()`` replaces the Python
wrappers of the CUDA entry points (``_device.*`` and the low-level functions of ``_stats_device``) with functions that
compute the same quantities through the CPU oracle (``oracle/c_oracle.py``, ``oracle/np_port.py``) and NumPy, on
``torch`` CPU tensors.  Nothing in the product imports this module; the kernels themselves are checked against the
same oracle by the ``-m gpu`` tests.
"""
from __future__ import annotations

import numpy as np
import torch

from mcframework_b200 import _device as D
from mcframework_b200 import _native as N
from mcframework_b200 import _stats_device as SD
from mcframework_b200 import _streams as S
from oracle import c_oracle as co
from oracle import np_port

_M64 = (1 << 64) - 1


def _gen_of(rec) -> co.Gen:
    """Oracle generator positioned where the ``mcf_stream`` record says the reference stream is."""
    s = [int(v) for v in rec["s"]]
    if int(rec["kind"]) == S.GEN_PCG64:
        st = {"state": {"state": (s[0] << 64) | s[1], "inc": (s[2] << 64) | s[3]}, "has_uint32": int(rec["has_uint32"]),
              "uinteger": int(rec["uinteger"])}
        return co.Gen.pcg64_from_state(st)
    if any(s[2:6]) or int(rec["buf_pos"]) != 4:
        raise NotImplementedError("fake device: Philox streams must be fresh (counter 0, empty buffer)")
    return co.Gen.philox_from_key(s[:2])


def _records(dl):
    return dl.layout.records


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ------------------------------------------------------------------------------------------- _device
def layout_of(cls, layout):
    return cls(layout, None)


def pi_hits(dl, n_points, antithetic=False):
    outs, hits = [], []
    for rec in _records(dl):
        o, h = _gen_of(rec).pi(int(rec["n_sims"]), int(n_points), bool(antithetic))
        outs.append(o)
        hits.append(h)
    return _t(np.concatenate(hits)), _t(np.concatenate(outs))


class _FakePlan(D.NormalPlan):
    def status(self) -> int:  # the stand-in plan is exact by construction
        return N.OK


def normal_plan(dl, normals_per_sim, slack=None, resolve_iters=3, max_retries=4, check=True):
    used = []
    for rec in _records(dl):
        g = _gen_of(rec)
        g.standard_normal(int(rec["n_sims"]) * int(normals_per_sim))
        used.append(g.consumed)
    return _FakePlan(None, _t(np.array(used, dtype=np.int64)), int(normals_per_sim), None, float(slack or 0.03))


def _terminal(g, n, n_steps, mode, params):
    if mode == N.NORMAL_SUM:
        return g.standard_normal(n * n_steps).reshape(n, n_steps).sum(axis=1)
    if mode == N.NORMAL_LOC_SCALE:
        z = g.standard_normal(n * n_steps).reshape(n, n_steps)
        return (float(params[0]) + float(params[1]) * z).sum(axis=1)
    return g.gbm_terminal(n, n_steps, mode, params)


def gbm_terminal(dl, n_steps, plan, specs):
    lay = dl.layout
    out = np.empty((len(specs), lay.n_total))
    for k, (mode, params) in enumerate(specs):
        for rec in _records(dl):
            f, n = int(rec["first_sim"]), int(rec["n_sims"])
            out[k, f:f + n] = _terminal(_gen_of(rec), n, int(n_steps), int(mode), list(params))
    return _t(out)


def gbm_paths(dl, n_steps, plan, S0, r, sigma, T, time_major=False):
    parts = [_gen_of(rec).gbm_paths(int(rec["n_sims"]), int(n_steps), S0, r, sigma, T) for rec in _records(dl)]
    paths = np.concatenate(parts, axis=0)
    return _t(paths.T.copy() if time_major else paths)


def draw_integers(dl, low, high, size, op, match=0):
    outs, rejected = [], 0
    for rec in _records(dl):
        n = int(rec["n_sims"])
        g = _gen_of(rec)
        vals = g.integers(int(low), int(high), n * int(size)).reshape(n, int(size))
        # the kernel's flag: does any of the first n*size 32-bit slots fail NumPy's Lemire test?
        span = int(high) - int(low)
        slots = _gen_of(rec).raw32(n * int(size)).astype(np.uint64)
        if np.any(((slots * np.uint64(span)) & np.uint64(0xFFFFFFFF)) < np.uint64(((1 << 32) - span) % span)):
            rejected = 1
        if int(op) == 0:
            outs.append(vals.sum(axis=1).astype(np.float64))
        else:
            runs = np.zeros(n)
            for i, row in enumerate(vals):
                best = cur = 0
                for v in row:
                    cur = cur + 1 if v == match else 0
                    best = max(best, cur)
                runs[i] = best
            outs.append(runs)
    return _t(np.concatenate(outs)), torch.tensor([rejected], dtype=torch.int32)


# ------------------------------------------------------------------------------------- _stats_device
def moments_raw(x, target=0.0):
    a = x.numpy().astype(np.float64, copy=False).reshape(-1)
    fin = a[np.isfinite(a)]
    out = np.zeros(SD.MOMENTS_OUT)
    out[0], out[1] = fin.size, np.isnan(a).sum()
    out[2], out[3] = np.isposinf(a).sum(), np.isneginf(a).sum()
    if fin.size:
        m = fin.mean()
        d = fin - m
        out[4], out[5], out[6], out[7] = m, (d ** 2).sum(), (d ** 3).sum(), (d ** 4).sum()
        out[8], out[9] = ((fin - float(target)) ** 2).sum(), fin.sum()
    return _t(out)


def select(x, ranks, omit_nonfinite=False, group=None):
    a = x.numpy().reshape(-1)
    if group is not None:  # sharded sample: the order statistics of the union (the kernels all-reduce digit histograms)
        import torch.distributed as dist

        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, a, group=group)
        a = np.concatenate(parts)
    if omit_nonfinite:
        a = a[np.isfinite(a)]
    return _t(np.sort(a)[[int(r) for r in ranks]].astype(np.float64))


def bootstrap_means(x, n_resamples, gen, group=None, advance=True, **_tuning):
    a = x.numpy().reshape(-1)
    n, B = a.size, int(n_resamples)
    if n == 1:
        return x.to(torch.float64).expand(B).clone()
    if n < 2 or n >= (1 << 32):
        raise ValueError("bootstrap_means needs 1 <= n < 2**32")
    saved = gen.bit_generator.state
    means = np_port.bootstrap_means_chunked(a, B, gen)
    if not advance:
        gen.bit_generator.state = saved
    return _t(np.asarray(means, dtype=np.float64))


def transpose(a):
    return a.t().contiguous()


def lsm_price(paths_tm, K, r, dt, is_put, group=None, return_exercise=False):
    if return_exercise:
        raise NotImplementedError("fake device: exercise steps are a device-only extra")
    price, _ = np_port.lsm_price(paths_tm.numpy().T, float(K), float(r), float(dt), "put" if is_put else "call")
    return torch.tensor([float(price)], dtype=torch.float64)


def count_less(x, value):
    return int((x.numpy() < float(value)).sum())


def compact_finite(x):
    a = x.numpy().reshape(-1)
    return _t(a[np.isfinite(a)].copy())


# --------------------------------------------------------------------------------------- installation
def _cpu_device(fn):
    def wrapped(*args, **kw):
        if kw.get("device") in ("cuda", torch.device("cuda")):
            kw["device"] = "cpu"
        return fn(*args, **kw)

    return wrapped


def install():
    """Route the device entry points to the oracle and every ``device="cuda"`` allocation to the CPU."""
    N.lib = lambda: object()  # "a device is present": nothing below dereferences it
    D.DeviceLayout.of = classmethod(layout_of)
    for name in ("pi_hits", "normal_plan", "gbm_terminal", "gbm_paths", "draw_integers"):
        setattr(D, name, globals()[name])
    for name in ("moments_raw", "select", "bootstrap_means", "transpose", "lsm_price", "count_less", "compact_finite"):
        setattr(SD, name, globals()[name])
    for name in ("empty", "zeros", "full", "tensor"):
        setattr(torch, name, _cpu_device(getattr(torch, name)))
    real_to = torch.Tensor.to

    def to(self, *args, **kw):
        args = tuple("cpu" if (isinstance(a, str) and a == "cuda") else a for a in args)
        if kw.get("device") in ("cuda", torch.device("cuda")):
            kw["device"] = "cpu"
        return real_to(self, *args, **kw)

    torch.Tensor.to = to
