"""ctypes loader for libqzb200.so (the sm_100a CUDA library behind include/qzb200.h).

There is NO fallback: if the library is missing, or a call fails, this module raises.  PyTorch is used
only as the owner of device memory and streams; every kernel launched here is ours.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_PKG, "libqzb200.so")

ABI_VERSION = 2
N_ACTIONS = 140
STATE_ELEMS = 26 * 9 * 9
DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1


class QzError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); mirrors include/qzb200.h (tests/test_capi_symbols.py cross-checks the header)
_vp, _i64, _i32, _u64, _f64, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_double, C.c_float
SIGNATURES = {
    "qz_version": (C.c_int, []),
    "qz_last_error_string": (C.c_char_p, []),
    "qz_device_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "qz_stream_check": (C.c_int, [_vp]),
    "qz_env_reset": (C.c_int, [_vp, _i64, _vp]),
    "qz_env_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "qz_env_legal_mask": (C.c_int, [_vp, _vp, _i64, _vp]),
    "qz_env_legal_mask_flagged": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _vp, _i32, C.c_int, _vp]),
    "qz_env_encode": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _i64, _vp]),
    "qz_env_sample_legal": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _i64, _vp]),
    "qz_env_random_play": (C.c_int, [_vp, _u64, _vp, _i32, _i64, _vp]),
    "qz_rollout_workspace_bytes": (C.c_int64, [_i64]),
    "qz_rollout_pawn_passes": (C.c_int32, [C.c_int32]),
    "qz_rollout": (C.c_int, [_vp, _i64, _vp, _i32, _i64, _u64, _u64, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "qz_rollout_finish": (C.c_int, [_vp, _i64, _vp, _i32, _i64, _u64, _u64, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "qz_mcts_backup_pending": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "qz_mcts_init": (C.c_int, [_vp, _vp, _vp, _vp]),
    "qz_mcts_select": (C.c_int, [_vp, _f64, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "qz_mcts_extend": (C.c_int, [_vp, _f64, _vp, C.c_int, _vp, _vp]),
    "qz_mcts_expand_backup": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f64, C.c_int, _vp, _vp]),
    "qz_mcts_root_stats": (C.c_int, [_vp, _f64, _vp, _vp, _vp, _vp, _vp]),
    "qz_mcts_choose": (C.c_int, [_vp, C.c_int, _f64, _f64, _f64, _u64, _vp, _vp, _vp]),
    "qz_mcts_reroot": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp]),
    "qz_mcts_node_children": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "qz_stub_eval": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _i64, _vp]),
}


def load():
    """Load the library once; raise QzError (never fall back) if it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise QzError(
            "libqzb200.so not found at %s -- build it with `python -m alphazero_quoridor_b200.build` "
            "(there is no CPU fallback)" % SO_PATH)
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise QzError("libqzb200.so does not export %s -- rebuild it" % name)
        fn.restype = res
        fn.argtypes = args
    if lib.qz_version() != ABI_VERSION:
        raise QzError("libqzb200.so ABI %d != expected %d -- rebuild it" % (lib.qz_version(), ABI_VERSION))
    _lib = lib
    return lib


LAUNCHES = 0      # kernels of ours launched through the C ABI (bench.py reports it as gpu_launches)


# kernels launched by one successful call of each entry point; the rollout entry points pass their own count
# (wall + stuck + qz_rollout_pawn_passes(limit) pawn passes, see rollout.py)
KERNELS_PER_CALL = {"qz_env_random_play": 2, "qz_stream_check": 0}


def check(rc, what="", launches=None):
    global LAUNCHES
    LAUNCHES += KERNELS_PER_CALL.get(what, 1) if launches is None else launches
    if rc != 0:
        msg = load().qz_last_error_string().decode("utf-8", "replace")
        raise QzError("%s failed with code %d: %s" % (what or "libqzb200 call", rc, msg))


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Raw device pointer of a contiguous CUDA tensor (or NULL for None)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise QzError("expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise QzError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def require_cuda():
    if not torch.cuda.is_available():
        raise QzError("no CUDA device: alphazero_quoridor_b200 has no CPU fallback")


def c_void_p_of(t):
    """Raw device pointer without the contiguity check (channels_last tensors are dense but not
    `is_contiguous()`); the caller vouches for the layout."""
    if not t.is_cuda:
        raise QzError("expected a CUDA tensor (there is no CPU path)")
    return C.c_void_p(t.data_ptr())
