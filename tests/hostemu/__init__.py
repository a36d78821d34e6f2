"""CPU emulation harness of the device functions (tests only). See hostemu.cpp."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
#: extra compiler flags (e.g. "-DMCF_SCAN_DENORM") select a code-generation variant of the device headers; each set of
#: flags gets its own library
_FLAGS = os.environ.get("MCF_HOSTEMU_FLAGS", "").split()
_SO = os.path.join(_HERE, "_build", "libhostemu%s.so" % ("_" + "".join(c for c in "".join(_FLAGS) if c.isalnum()) if _FLAGS else ""))
_SRC = [os.path.join(_HERE, "hostemu.cpp")] + [
    os.path.join(_HERE, "..", "..", "mcframework_b200", "csrc", f) for f in ("mcf_device.cuh", "mcf_host.h", "zig_tables.h")
]


def build():
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRC):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math"] + _FLAGS +
                              ["-x", "c++", _SRC[0], "-o", _SO, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class GbmSpec(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 6)]


def specs_array(specs):
    arr = (GbmSpec * len(specs))()
    for i, (mode, p) in enumerate(specs):
        arr[i].mode = mode
        for k, v in enumerate(p):
            arr[i].p[k] = float(v)
    return arr


def pi_hits(layout, n_points, antithetic=False):
    hits = np.zeros(layout.n_total, dtype=np.int64)
    lib().emu_pi_hits(C.c_uint32(layout.kind), _vp(layout.records), layout.n_streams, C.c_int64(layout.n_total),
                      C.c_int64(n_points), int(bool(antithetic)), C.c_int64(layout.hint), _vp(hits))
    return hits


class Plan:
    """Host-emulated plan: replication starts, stream ends and the emit masks the replay reads."""

    def __init__(self, rc, sim_start, stream_end, regen, emit, entry, zsum, cps):
        self.rc, self.sim_start, self.stream_end, self.regen = rc, sim_start, stream_end, regen
        self.emit, self.entry, self.zsum, self.cps = emit, entry, zsum, cps

    def __iter__(self):  # (rc, sim_start, stream_end, regenerated) -- the tuple older tests unpack
        return iter((self.rc, self.sim_start, self.stream_end, self.regen))


def normal_plan(layout, nps, slack=0.03, iters=3):
    sim_start = np.full(layout.n_total, np.iinfo(np.uint64).max, dtype=np.uint64)
    stream_end = np.zeros(layout.n_streams, dtype=np.uint64)
    regen = C.c_int64(0)
    cps = C.c_int64(0)
    lib().emu_plan_chunks.restype = C.c_int64
    n_chunks = lib().emu_plan_chunks(layout.n_streams, C.c_int64(layout.max_sims), C.c_int64(nps), C.c_double(slack),
                                     C.byref(cps))
    emit = np.zeros(n_chunks, dtype=np.uint64)
    entry = np.zeros(n_chunks, dtype=np.uint8)
    zsum = np.zeros(n_chunks * 4, dtype=np.float64)  # MCF_SUBS sums per chunk
    rc = lib().emu_normal_plan(C.c_uint32(layout.kind), _vp(layout.records), layout.n_streams,
                               C.c_int64(layout.max_sims), C.c_int64(nps), C.c_double(slack), iters, _vp(sim_start),
                               _vp(stream_end), C.byref(regen), _vp(emit), _vp(entry), _vp(zsum))
    return Plan(rc, sim_start, stream_end, regen.value, emit, entry, zsum, cps.value)


def gbm_terminal(layout, n_steps, plan, specs, masked=True):
    """``plan``: a Plan (mask-driven replay, what the kernels do) or a bare sim_start array (state machine)."""
    out = np.empty((len(specs), layout.n_total))
    arr = specs_array(specs)
    if isinstance(plan, Plan) and masked:
        rc = lib().emu_gbm_terminal(C.c_uint32(layout.kind), _vp(layout.records), layout.n_streams,
                                    C.c_int64(layout.n_total), C.c_int64(n_steps), C.c_int64(layout.hint),
                                    _vp(plan.sim_start), arr, len(specs), _vp(out), _vp(plan.emit), _vp(plan.entry),
                                    _vp(plan.zsum), C.c_int64(plan.cps), _vp(plan.stream_end))
    else:
        starts = plan.sim_start if isinstance(plan, Plan) else plan
        rc = lib().emu_gbm_terminal(C.c_uint32(layout.kind), _vp(layout.records), layout.n_streams,
                                    C.c_int64(layout.n_total), C.c_int64(n_steps), C.c_int64(layout.hint), _vp(starts),
                                    arr, len(specs), _vp(out), None, None, None, C.c_int64(0), None)
    assert rc == 0, rc
    return out


def gbm_paths(layout, n_steps, sim_start, p4, time_major=False):
    shape = (n_steps + 1, layout.n_total) if time_major else (layout.n_total, n_steps + 1)
    out = np.empty(shape)
    p = np.asarray(p4, dtype=np.float64)
    lib().emu_gbm_paths(C.c_uint32(layout.kind), _vp(layout.records), layout.n_streams, C.c_int64(layout.n_total),
                        C.c_int64(n_steps), C.c_int64(layout.hint), _vp(sim_start), _vp(p), int(time_major), _vp(out))
    return out


def raw64(record, pos, n):
    out = np.empty(n, dtype=np.uint64)
    lib().emu_raw64(C.c_uint32(int(record["kind"][0])), _vp(record), C.c_uint64(pos), C.c_int64(n), _vp(out))
    return out


def normals(record, pos, n):
    out = np.empty(n)
    end = C.c_uint64(0)
    lib().emu_normals(C.c_uint32(int(record["kind"][0])), _vp(record), C.c_uint64(pos), C.c_int64(n), _vp(out),
                      C.byref(end))
    return out, end.value


def philox_folded(record, blk0, n_blocks):
    """4*n_blocks raw draws from block ``blk0`` (relative to the record's counter) via philox_block_folded."""
    out = np.empty(4 * n_blocks, dtype=np.uint64)
    rc = lib().emu_philox_folded(_vp(record), C.c_uint64(blk0), C.c_int64(n_blocks), _vp(out))
    return rc, out
