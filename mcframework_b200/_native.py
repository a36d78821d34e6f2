"""ctypes binding of ``libmcf_b200.so`` (the C ABI declared in ``include/mcf_b200.h``).

There is NO CPU fallback: every device operation goes through :func:`lib`, which raises
``RuntimeError`` when the CUDA library is not built or no CUDA device is visible.
PyTorch is used only for device-memory ownership and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
#: developer override for A/B runs of alternative builds of the same ABI (tools/ab_variants.sh); unset in normal use
LIB_PATH = os.environ.get("MCF_B200_LIB") or os.path.join(_HERE, "libmcf_b200.so")

OK, ERR_BAD_ARG, ERR_CUDA, ERR_WORKSPACE_TOO_SMALL, ERR_WINDOW_TOO_SMALL, ERR_NOT_CONVERGED = 0, -1, -2, -3, -4, -5
_ERR_NAMES = {
    ERR_BAD_ARG: "MCF_ERR_BAD_ARG",
    ERR_CUDA: "MCF_ERR_CUDA",
    ERR_WORKSPACE_TOO_SMALL: "MCF_ERR_WORKSPACE_TOO_SMALL",
    ERR_WINDOW_TOO_SMALL: "MCF_ERR_WINDOW_TOO_SMALL",
    ERR_NOT_CONVERGED: "MCF_ERR_NOT_CONVERGED",
}

# replication bodies (enum in include/mcf_b200.h)
BS_EURO_CALL, BS_EURO_PUT, BS_AMER_CALL, BS_AMER_PUT, PATH_TERMINAL, PORTFOLIO_GBM, PORTFOLIO_LOG1P, NORMAL_SUM, NORMAL_LOC_SCALE = range(9)


class DeviceUnavailableError(RuntimeError):
    """The CUDA library is not built or no CUDA device is visible.  Never swallowed: there is no CPU fallback."""


class NativeError(RuntimeError):
    def __init__(self, fn: str, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"{fn} failed: {_ERR_NAMES.get(code, code)}{(' -- ' + detail) if detail else ''}")


class GbmSpec(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 6)]


_i32, _i64, _u32, _u64, _f64, _vp = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double, C.c_void_p

#: name -> (restype, argtypes); every symbol include/mcf_b200.h declares
SIGNATURES = {
    "mcf_abi_version": (C.c_int, []),
    "mcf_build_info": (C.c_char_p, []),
    "mcf_last_cuda_error": (C.c_int, []),
    "mcf_pi_hits": (C.c_int, [_u32, _vp, _vp, _i32, _i64, _i64, _i32, _i64, _vp, _vp, _vp]),
    "mcf_normal_plan_workspace_bytes": (_i64, [_i32, _i64, _i64, _f64]),
    "mcf_normal_plan": (C.c_int, [_u32, _vp, _vp, _i32, _i64, _i64, _i64, _f64, _i32, _vp, _i64, _vp, _vp, _vp]),
    "mcf_workspace_status": (C.c_int, [_vp, _vp]),
    "mcf_normal_relocate": (C.c_int, [_vp, _i32, _i64, _i64, _f64, _i64, _vp, _vp, _vp]),
    "mcf_gbm_slices_from_sums": (C.c_int, [_vp, _i64, _i64, _i64, _i64, C.POINTER(GbmSpec), _i32, _i32, _vp, _vp]),
    "mcf_draw_integers": (C.c_int, [_u32, _vp, _i32, _i64, _i64, _i64, _i64, _i32, _i32, _i64, _vp, _vp, _vp]),
    "mcf_gbm_terminal": (C.c_int, [_u32, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, C.POINTER(GbmSpec), _i32, _vp, _vp]),
    "mcf_gbm_paths": (C.c_int, [_u32, _vp, _i32, _i64, _i64, _i64, _vp, _vp, C.POINTER(_f64), _i32, _vp, _vp]),
    "mcf_gbm_slices": (C.c_int, [_u32, _vp, _i32, _i64, _i64, _i64, _vp, _vp, C.POINTER(GbmSpec), _i64, _i32, _vp, _vp]),
    # Longstaff-Schwartz
    "mcf_lsm_workspace_bytes": (_i64, [_i64]),
    "mcf_lsm_begin": (C.c_int, [_vp, _i64, _i64, _f64, _f64, _f64, _i32, _vp, _i64, _vp]),
    "mcf_lsm_reduce": (C.c_int, [_vp, _i64, _i64, _f64, _f64, _f64, _i32, _vp, _vp]),
    "mcf_lsm_apply": (C.c_int, [_vp, _i64, _i64, _f64, _f64, _f64, _i32, _vp, _vp]),
    "mcf_lsm_finish": (C.c_int, [_i64, _f64, _f64, _vp, _vp, _vp, _vp, _vp]),
    "mcf_lsm": (C.c_int, [_vp, _i64, _i64, _f64, _f64, _f64, _i32, _vp, _i64, _vp, _vp, _vp]),
    "mcf_select_bracket_bytes": (_i64, [_i64]),
    "mcf_select_bracketed": (C.c_int, [_vp, _i64, C.POINTER(C.c_int64), _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "mcf_probe_issue_rate": (C.c_int, [_i32, C.POINTER(C.c_double), _vp]),
    "mcf_lsm_exchange_bytes": (_i64, []),
    "mcf_peer_buffer_create": (C.c_int, [_i64, C.POINTER(C.c_void_p), _vp]),
    "mcf_peer_buffer_open": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "mcf_peer_buffer_close": (C.c_int, [_vp]),
    "mcf_peer_buffer_destroy": (C.c_int, [_vp]),
    "mcf_lsm_sharded": (C.c_int, [_vp, _i64, _i64, _i64, _f64, _f64, _f64, _i32, _vp, _i64, C.POINTER(C.c_void_p), _i32, _i32,
                                  C.c_uint64, _vp, _vp, _vp]),
    "mcf_transpose": (C.c_int, [_vp, _i64, _i64, _vp, _vp]),
    # StatsEngine reductions
    "mcf_moments_workspace_bytes": (_i64, [_i64]),
    "mcf_moments": (C.c_int, [_vp, _i64, _f64, _vp, _vp, _i64, _vp]),
    "mcf_select_workspace_bytes": (_i64, []),
    "mcf_select_hist_offset": (_i64, []),
    "mcf_select_hist_bytes": (_i64, []),
    "mcf_select_passes": (_i32, []),
    "mcf_select_begin": (C.c_int, [C.POINTER(_i64), _i32, _vp, _i64, _vp]),
    "mcf_select_hist": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "mcf_select_choose": (C.c_int, [_vp, _vp]),
    "mcf_select_finish": (C.c_int, [_vp, _vp, _vp]),
    "mcf_select": (C.c_int, [_vp, _i64, _i32, C.POINTER(_i64), _i32, _vp, _vp, _i64, _vp]),
    "mcf_select_compact_bytes": (_i64, [_i64]),
    "mcf_select_compacting": (C.c_int, [_vp, _i64, _i32, C.POINTER(_i64), _i32, _vp, _vp, _i64, _vp, _i64, _vp]),
    "mcf_set_l2_fetch_granularity": (C.c_int, [_i32]),
    "mcf_bootstrap_workspace_bytes": (_i64, [_i64, _i32]),
    "mcf_bootstrap_count": (C.c_int, [_u32, _vp, _i64, _u64, _u64, _i32, _vp, _i64, _vp, _vp]),
    "mcf_bootstrap_accumulate": (C.c_int, [_u32, _vp, _vp, _i64, _i64, _u64, _u64, _i32, _u64, _vp, _i64, _vp, _vp, _vp]),
    "mcf_bootstrap_finalize": (C.c_int, [_vp, _i64, _i64, _vp]),
    "mcf_bootstrap_blocks": (_i64, [_i64, _i32]),
    "mcf_bootstrap_block_base": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "mcf_bootstrap_binned_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "mcf_bootstrap_accumulate_binned": (C.c_int, [_u32, _vp, _vp, _i64, _i64, _u64, _u64, _i32, _u64, _vp, _i64, _vp, _i32, _i32,
                                                  _vp, _i64, _vp, _vp, _vp]),
    "mcf_bootstrap_fused_phase1": (C.c_int, [_u32, _vp, _vp, _i64, _i64, C.c_uint64, C.c_uint64, C.c_uint64, _i32, _vp, _i64, _i32,
                                             _i32, _vp, _i64, _vp, _vp, _vp]),
    "mcf_bootstrap_fused_phase2": (C.c_int, [_u32, _vp, _vp, _i64, _i64, C.c_uint64, C.c_uint64, C.c_uint64, _i32, C.c_uint64, _vp,
                                             _i64, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mcf_count_less": (C.c_int, [_vp, _i64, _f64, _vp, _vp]),
    "mcf_compact_workspace_bytes": (_i64, [_i64]),
    "mcf_compact_finite": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp]),
    "mcf_stats_last_cuda_error": (C.c_int, []),
}

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library and attach signatures (no CUDA call is made)."""
    if not os.path.exists(path):
        raise DeviceUnavailableError(
            f"CUDA library not built: {path} is missing. Run `python -m mcframework_b200.build` "
            "(needs nvcc; targets sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def lib():
    """The loaded library, after checking that a CUDA device is usable.  Fails loudly otherwise."""
    global _lib
    if _lib is None:
        import torch

        if not torch.cuda.is_available():
            raise DeviceUnavailableError("mcframework_b200 needs a CUDA device (B200, sm_100a); none is visible. "
                               "There is no CPU fallback.")
        _lib = load_library()
    return _lib


def check(fn: str, code: int):
    if code != OK:
        detail = ""
        if code == ERR_CUDA and _lib is not None:
            detail = f"cudaError {_lib.mcf_last_cuda_error() or _lib.mcf_stats_last_cuda_error()}"
        raise NativeError(fn, code, detail)
