"""Device-side operations behind the public API: thin Python over the C ABI.

All functions take/return ``torch`` CUDA tensors (ownership only) and launch on torch's current
stream.  Nothing here computes on the CPU; a missing library or device raises
(see :mod:`mcframework_b200._native`).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N
from ._streams import StreamLayout

#: counters for bench.py's ``gpu_launches`` (kernels launched by this library, by entry point)
LAUNCHES = {"pi": 0, "plan": 0, "gbm": 0, "paths": 0, "stats": 0, "lsm": 0}


KEY_BATCH = 112  # streams per launch of the launch-constant-key kernels (MCF_KEY_BATCH in csrc/mcf_kernels.cu)


def _batches(lay) -> int:
    """Launches one pass over the layout's streams takes: Philox streams go 112 per launch, anything else in one."""
    return -(-lay.n_streams // KEY_BATCH) if lay.kind == 1 else 1


def _torch():
    import torch

    return torch


def _stream_ptr():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def upload_records(layout: StreamLayout):
    """Copy the ``mcf_stream`` records to the device (a few KB)."""
    torch = _torch()
    host = torch.from_numpy(np.ascontiguousarray(layout.records).view(np.uint8).copy())
    return host.to("cuda", non_blocking=False)


@dataclass
class DeviceLayout:
    layout: StreamLayout
    records: "object"  # torch uint8 tensor on device

    @classmethod
    def of(cls, layout: StreamLayout) -> "DeviceLayout":
        N.lib()
        return cls(layout, upload_records(layout))


def pi_hits(dl: DeviceLayout, n_points: int, antithetic: bool = False):
    """(hits int64[n], estimates float64[n]) on device -- sims/pi.py:69-82 under the replication loop."""
    torch = _torch()
    lay = dl.layout
    hits = torch.empty(lay.n_total, dtype=torch.int64, device="cuda")
    out = torch.empty(lay.n_total, dtype=torch.float64, device="cuda")
    host_records = np.ascontiguousarray(lay.records)
    rc = N.lib().mcf_pi_hits(lay.kind, _ptr(dl.records), host_records.ctypes.data_as(C.c_void_p), lay.n_streams, lay.n_total,
                             int(n_points), int(bool(antithetic)), lay.hint, _ptr(hits), _ptr(out), _stream_ptr())
    N.check("mcf_pi_hits", rc)
    LAUNCHES["pi"] += 1 + _batches(lay)
    return hits, out


@dataclass
class NormalPlan:
    """Where every replication's first ``standard_normal`` draw sits in its reference stream."""

    sim_start: "object"  # torch int64 (bit pattern of uint64) [n_total]
    stream_end: "object"  # torch int64 [n_streams]: 64-bit draws consumed per stream
    normals_per_sim: int
    workspace: "object" = None  # torch uint8: emit masks etc. -- the replay kernels read it
    slack: float = 0.03  # window slack the workspace geometry was computed with (mcf_normal_relocate needs it)
    resolve_iters: int = 3

    def status(self) -> int:
        """Status of the planning pass (synchronises the stream: reads the workspace header back)."""
        return int(N.lib().mcf_workspace_status(_ptr(self.workspace), _stream_ptr()))


#: NumPy's ziggurat consumes 1.02201 64-bit draws per normal on average (measured over 2e8 normals; variance of the
#: count per normal ~0.04): the scan window of a stream is its expected consumption plus 8 standard deviations
ZIG_DRAWS_PER_NORMAL = 1.02201


def default_slack(normals_per_stream: int) -> float:
    need = max(1.0, float(normals_per_stream))
    return (ZIG_DRAWS_PER_NORMAL - 1.0) + 0.0004 + 8.0 * (0.04 / need) ** 0.5


def with_plan(dl: DeviceLayout, normals_per_sim: int, body, max_retries: int = 4):
    """``plan = normal_plan(...); out = body(plan)`` with ONE host round trip: the planning kernels and whatever ``body``
    launches are queued back to back, and the plan's status is read afterwards (a replay on a failed plan returns at once
    on the device).  A failed plan is re-planned with a wider window / more fix-up rounds and ``body`` runs again.
    Returns ``(plan, out)``."""
    slack, resolve_iters, status = None, 3, N.OK
    for _ in range(max_retries):
        plan = normal_plan(dl, normals_per_sim, slack, resolve_iters, check=False)
        out = body(plan)
        status = plan.status()
        if status == N.OK:
            return plan, out
        slack, resolve_iters = _replan_parameters(status, plan.slack, plan.resolve_iters)
        del plan, out
    raise N.NativeError("mcf_normal_plan", status, "planning did not succeed after retries")


def _replan_parameters(status: int, slack: float, resolve_iters: int):
    if status == N.ERR_WINDOW_TOO_SMALL:
        return slack * 2 + 0.02, resolve_iters
    if status == N.ERR_NOT_CONVERGED:
        return slack, min(resolve_iters + 3, 12)
    N.check("mcf_workspace_status", status)
    return slack, resolve_iters


def normal_plan(dl: DeviceLayout, normals_per_sim: int, slack: float | None = None, resolve_iters: int = 3,
                max_retries: int = 4, check: bool = True, buffers: dict | None = None) -> NormalPlan:
    """Planning pass for ziggurat-consuming replications ("stream-exact planning").
    ``check=False``: one attempt, no host synchronisation -- the caller reads ``plan.status()`` later (``with_plan``)."""
    torch = _torch()
    lay = dl.layout
    lib = N.lib()
    if slack is None:  # a window that turns out too small is re-planned with a wider one below
        slack = default_slack(int(lay.max_sims) * int(normals_per_sim))
    # ``buffers``: optional dict a caller keeps across calls ({"sim_start", "workspace"} tensors are reused when they are
    # large enough) -- a repeated plan of the same shape then allocates nothing
    sim_start = _reuse(buffers, "sim_start", lay.n_total, torch.int64)
    stream_end = torch.zeros(lay.n_streams, dtype=torch.int64, device="cuda")
    for _ in range(max_retries):
        nbytes = lib.mcf_normal_plan_workspace_bytes(lay.n_streams, lay.max_sims, int(normals_per_sim), float(slack))
        if nbytes < 0:
            N.check("mcf_normal_plan_workspace_bytes", int(nbytes))
        ws = _reuse(buffers, "workspace", int(nbytes), torch.uint8)
        host_records = np.ascontiguousarray(lay.records)
        rc = lib.mcf_normal_plan(lay.kind, _ptr(dl.records), host_records.ctypes.data_as(C.c_void_p), lay.n_streams,
                                 lay.n_total, lay.max_sims,
                                 int(normals_per_sim), float(slack), int(resolve_iters), _ptr(ws), int(nbytes),
                                 _ptr(sim_start), _ptr(stream_end), _stream_ptr())
        N.check("mcf_normal_plan", rc)
        LAUNCHES["plan"] += 4 + 2 * resolve_iters + _batches(lay)
        plan = NormalPlan(sim_start, stream_end, int(normals_per_sim), ws, float(slack), int(resolve_iters))
        if not check:
            return plan
        status = plan.status()
        if status == N.OK:
            return plan
        del ws, plan
        slack, resolve_iters = _replan_parameters(status, slack, resolve_iters)
    raise N.NativeError("mcf_normal_plan", status, "planning did not succeed after retries")


def _spec_array(specs):
    arr = (N.GbmSpec * len(specs))()
    for i, (mode, params) in enumerate(specs):
        arr[i].mode = int(mode)
        for k, v in enumerate(params):
            arr[i].p[k] = float(v)
    return arr


def _reuse(buffers, key: str, numel: int, dtype):
    """A tensor of exactly ``numel`` elements: a view of ``buffers[key]`` when that exists and is large enough, else fresh
    (and remembered in ``buffers`` for the next call)."""
    torch = _torch()
    if buffers is not None:
        have = buffers.get(key)
        if have is not None and have.dtype == dtype and have.numel() >= numel:
            return have[:numel]
    t = torch.empty(numel, dtype=dtype, device="cuda")
    if buffers is not None:
        buffers[key] = t
    return t


def gbm_terminal(dl: DeviceLayout, n_steps: int, plan: NormalPlan, specs, out=None):
    """float64[(len(specs), n)] terminal values; ``specs`` = [(mode, params), ...] on common random numbers.
    ``out``: optional float64 CUDA tensor of that shape to write into."""
    torch = _torch()
    lay = dl.layout
    if out is None:
        out = torch.empty((len(specs), lay.n_total), dtype=torch.float64, device="cuda")
    elif tuple(out.shape) != (len(specs), lay.n_total) or out.dtype != torch.float64 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float64 tensor of shape (len(specs), n)")
    host_records = np.ascontiguousarray(lay.records)
    rc = N.lib().mcf_gbm_terminal(lay.kind, _ptr(dl.records), host_records.ctypes.data_as(C.c_void_p), lay.n_streams,
                                  lay.n_total, int(n_steps), lay.hint, _ptr(plan.sim_start), _ptr(plan.workspace),
                                  _spec_array(specs), len(specs), _ptr(out), _stream_ptr())
    N.check("mcf_gbm_terminal", rc)
    affine = all(int(m) not in (N.BS_AMER_CALL, N.BS_AMER_PUT, N.PORTFOLIO_LOG1P) for m, _ in specs)
    per_stream = lay.kind == 1 and affine and int(n_steps) >= 16
    LAUNCHES["gbm"] += _batches(lay) if per_stream else 1  # boundary-form replay: one launch per batch of streams
    return out


def draw_integers(dl: DeviceLayout, low: int, high: int, size: int, op: int, match: int = 0):
    """Per replication ``rng.integers(low, high, size=size)`` reduced on the device (op 0: sum, op 1: longest run of
    ``match``).  Returns (float64[n] results, rejected flag tensor int32[1])."""
    torch = _torch()
    lay = dl.layout
    out = torch.empty(lay.n_total, dtype=torch.float64, device="cuda")
    rejected = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = N.lib().mcf_draw_integers(lay.kind, _ptr(dl.records), lay.n_streams, lay.n_total, lay.hint, int(low), int(high),
                                   int(size), int(op), int(match), _ptr(out), _ptr(rejected), _stream_ptr())
    N.check("mcf_draw_integers", rc)
    LAUNCHES["pi"] += 1
    return out, rejected


def gbm_paths(dl: DeviceLayout, n_steps: int, plan: NormalPlan, S0: float, r: float, sigma: float, T: float,
              time_major: bool = False):
    torch = _torch()
    lay = dl.layout
    shape = (n_steps + 1, lay.n_total) if time_major else (lay.n_total, n_steps + 1)
    out = torch.empty(shape, dtype=torch.float64, device="cuda")
    p4 = (C.c_double * 4)(float(S0), float(r), float(sigma), float(T))
    rc = N.lib().mcf_gbm_paths(lay.kind, _ptr(dl.records), lay.n_streams, lay.n_total, int(n_steps), lay.hint,
                               _ptr(plan.sim_start), _ptr(plan.workspace), p4, int(bool(time_major)), _ptr(out), _stream_ptr())
    N.check("mcf_gbm_paths", rc)
    LAUNCHES["paths"] += 1
    return out


def gbm_slices(dl: DeviceLayout, n_steps: int, plan: NormalPlan, mode: int, params, slice_stride: int,
               time_major: bool = True, dtype=None, timings: dict | None = None):
    """Path values at steps 0, stride, 2*stride, ...: (n_slices, n) when time_major else (n, n_slices);
    float64, or float32 with ``dtype=torch.float32`` (segment form only: the values are computed in fp64 and narrowed
    by the store).  ``timings``: optional dict that receives CUDA events around the three stages of the segment form."""
    torch = _torch()
    lay = dl.layout
    dtype = dtype or torch.float64
    if dtype not in (torch.float64, torch.float32):
        raise ValueError("slices are stored as float64 or float32")
    n_slices = int(n_steps) // int(slice_stride) + 1
    shape = (n_slices, lay.n_total) if time_major else (lay.n_total, n_slices)
    k = int(n_steps) // int(slice_stride)
    segment_form = (lay.kind == 1 and int(slice_stride) >= 16 and k * int(slice_stride) == int(n_steps)
                    and k > 1 and int(mode) in (N.PORTFOLIO_GBM, N.PATH_TERMINAL)
                    and bool((lay.records["buf_pos"] % 4 == 0).all()))
    out = torch.empty(shape, dtype=dtype if segment_form else torch.float64, device="cuda")

    def mark(name):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            timings[name] = e

    if segment_form:
        # Segment form: every path is cut into k pseudo-replications of `stride` steps.  Re-locating the finished plan
        # gives their stream positions, the boundary-form replay their normal sums (two Philox blocks per segment
        # instead of stride/4), and one accumulation pass the slice values.
        mark("start")
        recs = lay.records.copy()
        recs["first_sim"] *= k
        recs["n_sims"] *= k
        lay2 = StreamLayout(lay.kind, recs, lay.n_total * k, lay.hint * k)
        dl2 = DeviceLayout(lay2, upload_records(lay2))
        starts = torch.empty(lay2.n_total, dtype=torch.int64, device="cuda")
        rc = N.lib().mcf_normal_relocate(_ptr(dl2.records), lay.n_streams, lay.max_sims, int(plan.normals_per_sim),
                                         float(plan.slack), int(slice_stride), _ptr(plan.workspace), _ptr(starts), _stream_ptr())
        N.check("mcf_normal_relocate", rc)
        LAUNCHES["plan"] += 1
        mark("relocated")
        plan2 = NormalPlan(starts, plan.stream_end, int(slice_stride), plan.workspace, plan.slack)
        seg = gbm_terminal(dl2, int(slice_stride), plan2, [(N.NORMAL_SUM, [])])[0]
        del starts, plan2
        mark("segment_sums")
        rc = N.lib().mcf_gbm_slices_from_sums(_ptr(seg), lay.n_total, k, int(slice_stride), int(n_steps),
                                              _spec_array([(mode, params)]), int(bool(time_major)),
                                              int(dtype == torch.float32), _ptr(out), _stream_ptr())
        N.check("mcf_gbm_slices_from_sums", rc)
        LAUNCHES["paths"] += 1
        mark("stored")
        return out
    rc = N.lib().mcf_gbm_slices(lay.kind, _ptr(dl.records), lay.n_streams, lay.n_total, int(n_steps), lay.hint,
                                _ptr(plan.sim_start), _ptr(plan.workspace), _spec_array([(mode, params)]), int(slice_stride),
                                int(bool(time_major)), _ptr(out), _stream_ptr())
    N.check("mcf_gbm_slices", rc)
    LAUNCHES["paths"] += 1
    return out if dtype == torch.float64 else out.to(dtype)
