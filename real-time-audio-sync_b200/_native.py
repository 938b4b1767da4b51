"""ctypes binding of libafsync.so (include/afsync.h) + small torch helpers.

There is deliberately NO CPU fallback: importing this module without the built
library, or calling into it without a CUDA device, raises immediately.
PyTorch is used only as plumbing (device allocations, streams).
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# AFSYNC_LIB: developer override to time an alternative build of the same ABI (default: the in-tree library)
LIB_PATH = os.environ.get("AFSYNC_LIB") or os.path.join(_HERE, "libafsync.so")

AFS_F64, AFS_F32, AFS_BF16X3 = 0, 1, 2
AFS_OTW, AFS_LIVENOTE_V2, AFS_LIVENOTE_V1 = 0, 1, 2
AFS_COST_COSINE, AFS_COST_EUCLID = 0, 1
AFS_STEP_NONE, AFS_STEP_STOP, AFS_STEP_FULL = 0, 1, 2


class AfsError(RuntimeError):
    pass


_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p

_SIGNATURES = {
    # name: (restype, argtypes)
    "afs_last_error": (C.c_char_p, []),
    "afs_version": (C.c_char_p, []),
    "afs_launch_count": (C.c_int64, []),
    "afs_dtw_plan_create": (C.c_int, [C.POINTER(_vp), C.c_int, _i64p, _i64p, _i64p, _i64p, C.c_int, C.c_int]),
    "afs_dtw_plan_destroy": (C.c_int, [_vp]),
    "afs_dtw_plan_workspace_bytes": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "afs_dtw_plan_path_layout": (C.c_int, [_vp, C.c_int, _i64p, _i64p]),
    "afs_dtw_accumulate": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "afs_dtw_backtrack": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "afs_dtw_accumulate_stripe": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "afs_dtw_backtrack_stripe": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "afs_ipc_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp), _vp]),
    "afs_ipc_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "afs_ipc_clear": (C.c_int, [_vp, C.c_size_t, _vp]),
    "afs_ipc_close": (C.c_int, [_vp]),
    "afs_ipc_free": (C.c_int, [_vp]),
    "afs_otw_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, _i64p, _i64p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "afs_otw_destroy": (C.c_int, [_vp]),
    "afs_otw_state_bytes": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "afs_otw_reset": (C.c_int, [_vp, _vp, _vp]),
    "afs_otw_points_per_step": (C.c_int, [_vp]),
    "afs_otw_seed_set_live": (C.c_int, [_vp, _vp]),
    "afs_otw_step": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "afs_otw_path_layout": (C.c_int, [_vp, C.c_int, _i64p, _i64p]),
    "afs_otw_path_ptr": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "afs_otw_positions_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "afs_otw_read_window": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "afs_chroma_plan_create": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int, C.c_int]),
    "afs_chroma_plan_destroy": (C.c_int, [_vp]),
    "afs_chroma_num_frames": (C.c_int64, [_vp, C.c_int64, C.c_int]),
    "afs_chroma_batch": (C.c_int, [_vp, _vp, _i64p, C.c_int, C.c_int, C.c_int, _vp, _i64p, C.c_int, C.c_int, _vp]),
    "afs_chroma_batch_pcm16": (C.c_int, [_vp, _vp, _i64p, C.c_int, C.c_int, C.c_int, _vp, _i64p, C.c_int, C.c_int, _vp]),
    "afs_stft_batch": (C.c_int, [_vp, _vp, _i64p, C.c_int, C.c_int, _vp, _i64p, C.c_int, _vp]),
    "afs_wtw_create": (C.c_int, [C.POINTER(_vp), C.c_int, _vp, _i64p, _i64p, C.c_int, C.c_int, C.c_int]),
    "afs_wtw_destroy": (C.c_int, [_vp]),
    "afs_wtw_state_bytes": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "afs_wtw_reset": (C.c_int, [_vp, _vp, _vp]),
    "afs_wtw_push": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "afs_wtw_push_audio": (C.c_int, [_vp, _vp, _vp, _i64p, C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "afs_wtw_path_layout": (C.c_int, [_vp, C.c_int, _i64p, _i64p]),
    "afs_wtw_path_ptr": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "afs_wtw_positions_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
}

_lib = None


def exported_symbols():
    """Names every entry point include/afsync.h declares (checked by the CPU tests)."""
    return sorted(_SIGNATURES)


def lib():
    """Load libafsync.so (no compute happens here, so this also works without a GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AfsError(
                "libafsync.so is not built (%s missing). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python real-time-audio-sync_b200/build.py`. There is no CPU fallback." % LIB_PATH
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise AfsError("libafsync error %d: %s" % (status, lib().afs_last_error().decode("utf-8", "replace")))


def require_cuda():
    if not torch.cuda.is_available():
        raise AfsError("no CUDA device visible: the alignment hot path runs on the GPU only (no CPU fallback)")


def device(index=None):
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device() if index is None else index)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) address of a tensor as c_void_p; None -> NULL."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def i64_array(values):
    arr = np.ascontiguousarray(values, dtype=np.int64)
    return arr, arr.ctypes.data_as(_i64p)


def launch_count():
    return int(lib().afs_launch_count())
