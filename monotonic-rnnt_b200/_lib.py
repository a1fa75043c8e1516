"""ctypes binding of libmonotonic_rnnt.so -- the flat C ABI of include/mrnnt_c_api.h.

This is the ONLY way the Python host code reaches the kernels.  There is no Python / CPU fallback:
if the library is missing, cannot be built or cannot be loaded, importing callers get a loud error.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

STATUS_TEXT = {0: "no error", 1: "cuda memcpy or memset failed", 2: "invalid value", 3: "execution failed",
               4: "unknown error"}

OPT_FORCE_GENERIC = 1
OPT_TIMING = 2
OPT_K1_WARPS = 3
OPT_K3_WARPS = 4
OPT_K2_PARTS = 5
OPT_RESERVED_SMS = 6
OPT_PDL = 7
OPT_K1_COMPACT = 8
OPT_K2_ZERO_FILL = 9
OPT_DYNAMIC_TILES = 10
OPT_LAUNCH_COUNT = 11
OPT_UPLOAD_COPY_ENGINE = 12
OPT_RETURN_EARLY = 13
OPT_K2_FILL_SHARE = 14
OPT_FUSED_PLAN = 15
DBG_DENOM, DBG_ALPHA, DBG_BETA, DBG_LP, DBG_BAND, DBG_ROWMETA, DBG_LL, DBG_ROWSTART = range(1, 9)

# every symbol include/mrnnt_c_api.h and include/rnnt_entrypoint.h declare
EXPORTED_SYMBOLS = (
    "compute_rnnt_loss", "mrnnt_get_workspace_size", "mrnnt_create", "mrnnt_destroy", "mrnnt_workspace_size",
    "mrnnt_set_workspace", "mrnnt_create_workspace", "mrnnt_free_workspace", "mrnnt_restrict_to_alignment",
    "mrnnt_cost_and_grad", "mrnnt_enqueue", "mrnnt_device_costs", "rnnt_loss_grad_gpu", "mrnnt_set_option", "mrnnt_get_option",
    "mrnnt_debug_copy", "mrnnt_synth_uniform", "mrnnt_build_info", "mrnnt_last_timings",
    "mrnnt_enqueue_forward", "mrnnt_enqueue_backward", "mrnnt_enqueue_forward_into", "mrnnt_create_padded",
    "mrnnt_get_workspace_size_padded", "mrnnt_set_dtype",
    "mrnnt_peer_board_create", "mrnnt_peer_board_open", "mrnnt_peer_board_close", "mrnnt_peer_board_destroy",
    "mrnnt_set_peer_reduce", "mrnnt_peer_epoch", "mrnnt_upload_acts",
    "get_workspace_size", "mrnnt_restrict_to_alignment_strided", "mrnnt_get_shape", "mrnnt_set_workspace_cache_limit",
    "mrnnt_trim_workspace_cache", "mrnnt_set_peer_timeout_ms", "mrnnt_peer_failed",
)


class RNNTError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = int(status)
        super().__init__(f"{where}: status {status} ({STATUS_TEXT.get(int(status), '?')})")


_lib = None


def _declare(lib: ctypes.CDLL) -> None:
    vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    szp = ctypes.POINTER(ctypes.c_size_t)
    lib.mrnnt_get_workspace_size.argtypes = [vp, vp, ci, ci, szp]
    lib.mrnnt_create.argtypes = [ctypes.POINTER(vp), vp, vp, ci, vp, vp, ci, vp, vp]
    lib.mrnnt_get_workspace_size_padded.argtypes = [vp, vp, ci, ci, ci, ci, ci, szp]
    lib.mrnnt_create_padded.argtypes = [ctypes.POINTER(vp), vp, vp, ci, vp, vp, ci, ci, ci, ci, vp, vp]
    lib.mrnnt_set_dtype.argtypes = [vp, ci]
    lib.mrnnt_destroy.argtypes = [vp]
    lib.mrnnt_destroy.restype = None
    lib.mrnnt_workspace_size.argtypes = [vp, szp]
    lib.mrnnt_set_workspace.argtypes = [vp, vp]
    lib.mrnnt_create_workspace.argtypes = [vp]
    lib.mrnnt_free_workspace.argtypes = [vp]
    lib.mrnnt_free_workspace.restype = None
    lib.mrnnt_restrict_to_alignment.argtypes = [vp, vp, ci, ci]
    lib.mrnnt_restrict_to_alignment_strided.argtypes = [vp, vp, ci, ci, ci]
    lib.mrnnt_restrict_to_alignment_strided.restype = ci
    lib.mrnnt_get_shape.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ctypes.c_int64)]
    lib.mrnnt_get_shape.restype = ci
    lib.get_workspace_size.argtypes = [vp, vp, ci, ci, szp]
    lib.get_workspace_size.restype = ci
    lib.mrnnt_set_workspace_cache_limit.argtypes = [sz]
    lib.mrnnt_set_workspace_cache_limit.restype = None
    lib.mrnnt_trim_workspace_cache.argtypes = []
    lib.mrnnt_trim_workspace_cache.restype = None
    lib.mrnnt_set_peer_timeout_ms.argtypes = [vp, ctypes.c_uint]
    lib.mrnnt_set_peer_timeout_ms.restype = ci
    lib.mrnnt_peer_failed.argtypes = [vp]
    lib.mrnnt_peer_failed.restype = ci
    lib.mrnnt_upload_acts.argtypes = [vp, vp, vp]
    lib.mrnnt_upload_acts.restype = ci
    lib.mrnnt_cost_and_grad.argtypes = [vp, ci, vp, vp, vp]
    lib.mrnnt_enqueue.argtypes = [vp, ci, vp, vp]
    lib.mrnnt_enqueue_forward.argtypes = [vp, ci, vp, ci]
    lib.mrnnt_enqueue_backward.argtypes = [vp, vp, vp, vp]
    lib.mrnnt_enqueue_forward_into.argtypes = [vp, ci, vp, vp]
    lib.mrnnt_device_costs.argtypes = [vp]
    lib.mrnnt_device_costs.restype = vp
    lib.rnnt_loss_grad_gpu.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, vp, ci, vp, sz, vp, vp, vp]
    lib.mrnnt_set_option.argtypes = [vp, ci, ci]
    lib.mrnnt_get_option.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_int)]
    lib.mrnnt_debug_copy.argtypes = [vp, ci, vp, sz]
    lib.mrnnt_last_timings.argtypes = [vp, vp]
    lib.mrnnt_synth_uniform.argtypes = [vp, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64, vp]
    lib.mrnnt_peer_board_create.argtypes = [ci, ctypes.POINTER(vp), vp]
    lib.mrnnt_peer_board_open.argtypes = [vp, ctypes.POINTER(vp)]
    lib.mrnnt_peer_board_close.argtypes = [vp]
    lib.mrnnt_peer_board_destroy.argtypes = [vp]
    lib.mrnnt_set_peer_reduce.argtypes = [vp, ci, ci, ctypes.POINTER(vp), vp, ctypes.c_uint]
    lib.mrnnt_peer_epoch.argtypes = [vp]
    lib.mrnnt_peer_epoch.restype = ctypes.c_uint
    for name in ("mrnnt_peer_board_create", "mrnnt_peer_board_open", "mrnnt_peer_board_close",
                 "mrnnt_peer_board_destroy", "mrnnt_set_peer_reduce"):
        getattr(lib, name).restype = ci
    lib.mrnnt_build_info.argtypes = []
    lib.mrnnt_build_info.restype = ctypes.c_char_p
    for name in ("mrnnt_get_workspace_size", "mrnnt_create", "mrnnt_workspace_size", "mrnnt_set_workspace",
                 "mrnnt_create_workspace", "mrnnt_restrict_to_alignment", "mrnnt_cost_and_grad", "mrnnt_enqueue",
                 "rnnt_loss_grad_gpu", "mrnnt_set_option", "mrnnt_get_option", "mrnnt_debug_copy", "mrnnt_synth_uniform",
                 "mrnnt_last_timings", "mrnnt_enqueue_forward", "mrnnt_enqueue_backward", "mrnnt_enqueue_forward_into", "mrnnt_create_padded",
                 "mrnnt_get_workspace_size_padded", "mrnnt_set_dtype"):
        getattr(lib, name).restype = ci


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building first if the sources are newer and nvcc is available) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not _build.up_to_date():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this machine: a prebuilt library is acceptable, nothing else is
            if not os.path.exists(path):
                raise RuntimeError(
                    f"monotonic-rnnt_b200: {path} is missing and could not be built ({exc}); "
                    "there is no CPU fallback") from exc
    lib = ctypes.CDLL(path)
    _declare(lib)
    _lib = lib
    return lib


def check(status: int, where: str) -> None:
    if status != 0:
        raise RNNTError(status, where)
