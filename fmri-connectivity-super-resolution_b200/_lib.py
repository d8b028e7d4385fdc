"""ctypes binding of libgsr_sm100.so (include/gsr_sm100.h).

There is no CPU fallback: ``lib()`` raises if the shared library has not been
built (``python fmri-connectivity-super-resolution_b200/build.py``) and every
kernel wrapper raises on a non-CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgsr_sm100.so")
ABI_VERSION = 1

_lock = threading.Lock()
_lib = None

c_f32p = C.c_void_p      # device pointers travel as integers (tensor.data_ptr())
c_ptr = C.c_void_p
i32, i64, f32 = C.c_int, C.c_longlong, C.c_float

# name -> (restype, argtypes); must list EVERY symbol declared in include/gsr_sm100.h
PROTOTYPES = {
    "gsr_abi_version": (i32, []),
    "gsr_last_error": (C.c_char_p, []),
    "gsr_launch_count": (C.c_uint64, []),
    "gsr_set_tuning": (i32, [C.c_char_p, i32]),
    "gsr_normalize_adj_fwd": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, c_ptr]),
    "gsr_normalize_adj_bwd": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, c_ptr]),
    "gsr_normalize_adj_bwd_workspace": (C.c_size_t, [i32, i32]),
    "gsr_normalize_adj_bwd_tiles": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, c_ptr]),
    "gsr_normalize_adj_bwd_tiles2": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, c_ptr]),
    "gsr_eigh_workspace": (C.c_size_t, [i32, i32]),
    "gsr_eigh_jacobi_f32": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, f32, i32, c_ptr]),
    "gsr_eigh_tridiag_workspace": (C.c_size_t, [i32, i32]),
    "gsr_eigh_tridiag_f32": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, i32, i32, i32, c_ptr]),
    "gsr_bgemm_f32": (i32, [c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, i32, i32, i64, i64, i64,
                            i32, i32, i32, f32, i32, c_ptr, i32, i64, c_ptr]),
    "gsr_bgemm_reduce_workspace": (C.c_size_t, [i32, i32, i32, C.POINTER(i32)]),
    "gsr_bgemm_reduce_f32": (i32, [c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, i32, i32, i64, i64,
                                   i32, i32, i32, f32, i32, c_ptr, i32, c_ptr, i32, c_ptr]),
    "gsr_tcgemm_supported": (i32, [i32, i32, i32, i32, i32]),
    "gsr_tcgemm_f32": (i32, [c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, i32, i32, i64, i64, i64,
                             i32, i32, i32, f32, i32, c_ptr, i32, i64, c_ptr]),
    "gsr_tcgemm_dual_f32": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                                  i64, i64, i64, i64, i64, i32, i32, i32, f32, c_ptr]),
    "gsr_tcgemm_reduce_workspace": (C.c_size_t, [i32, i32, i32, C.POINTER(i32)]),
    "gsr_tcgemm_reduce_f32": (i32, [c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, i32, i32, i64, i64,
                                    i32, i32, i32, f32, i32, c_ptr, i32, c_ptr, i32, c_ptr]),
    "gsr_colsum_f32": (i32, [c_ptr, c_ptr, i32, i32, i32, c_ptr]),
    "gsr_splitk_reduce_f32": (i32, [c_ptr, c_ptr, i32, i32, i32, i32, f32, i32, c_ptr, i32, c_ptr]),
    "gsr_sym_diag_abs_fwd": (i32, [c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_sym_diag_abs_bwd": (i32, [c_ptr, c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_sym_diag_abs_l1_partials": (i32, [i32]),
    "gsr_sym_diag_abs_l1_fwd": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_sym_diag_abs_l1_bwd": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_abs_diag1_fwd": (i32, [c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_pool_scores": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_pool_topk": (i32, [c_ptr, i32, i32, c_ptr, c_ptr, c_ptr]),
    "gsr_pool_gather_scale_fwd": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_pool_bwd": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, c_ptr]),
    "gsr_gather_rows_cols": (i32, [c_ptr, c_ptr, c_ptr, i32, i32, i32, c_ptr]),
    "gsr_unpool_fwd": (i32, [c_ptr, c_ptr, c_ptr, i32, i32, i32, c_ptr]),
    "gsr_unpool_bwd": (i32, [c_ptr, c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_l1_partials": (i32, [i64]),
    "gsr_l1_loss_fwd": (i32, [c_ptr, c_ptr, i64, c_ptr, c_ptr, i32, i64, c_ptr]),
    "gsr_l1_loss_bwd": (i32, [c_ptr, c_ptr, i64, c_ptr, c_ptr, i32, i64, c_ptr]),
    "gsr_l1_loss_bwd_shared_y": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, i32, i64, c_ptr]),
    "gsr_vectorize_len": (i64, [i32, i32]),
    "gsr_vectorize": (i32, [c_ptr, c_ptr, i32, i32, i32, i32, c_ptr]),
    "gsr_anti_vectorize": (i32, [c_ptr, c_ptr, i32, i32, i64, i32, c_ptr]),
    "gsr_triu_flatten": (i32, [c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_eval_workspace": (C.c_size_t, []),
    "gsr_eval_reduce": (i32, [c_ptr, c_ptr, i64, i32, i32, c_ptr, c_ptr, c_ptr]),
    "gsr_flat_pack": (i32, [c_ptr, c_ptr, i32, c_ptr, i64, f32, c_ptr]),
    "gsr_flat_unpack": (i32, [c_ptr, c_ptr, i32, c_ptr, i64, f32, c_ptr]),
    "gsr_ffma_probe": (C.c_double, [i32, c_ptr, c_ptr]),
    "gsr_tf32_probe": (C.c_double, [i32, c_ptr]),
    "gsr_adam_step": (i32, [c_ptr] * 5 + [i32, c_ptr, f32, f32, f32, f32, f32, c_ptr]),
    "gsr_bn_stats_f32": (i32, [c_ptr, i32, i32, c_ptr, c_ptr, c_ptr, c_ptr, f32, c_ptr]),
    "gsr_bn_act_fwd_f32": (i32, [c_ptr] * 5 + [f32, f32, C.c_ulonglong, f32, c_ptr, i32, i32, c_ptr]),
    "gsr_bn_act_bwd_f32": (i32, [c_ptr] * 6 + [f32, f32, C.c_ulonglong, f32, i32, c_ptr, c_ptr, c_ptr, i32, i32, c_ptr]),
    "gsr_add_act_fwd_f32": (i32, [c_ptr, c_ptr, c_ptr, i64, f32, c_ptr]),
    "gsr_act_bwd_f32": (i32, [c_ptr, c_ptr, c_ptr, i64, f32, c_ptr]),
    "gsr_sigmoid_fwd_f32": (i32, [c_ptr, c_ptr, i64, c_ptr]),
    "gsr_sigmoid_bwd_f32": (i32, [c_ptr, c_ptr, c_ptr, i64, c_ptr]),
    "gsr_scale_div_f32": (i32, [c_ptr, c_ptr, c_ptr, i64, c_ptr]),
    "gsr_spectral_bwd_workspace": (C.c_size_t, []),
    "gsr_spectral_bwd_f32": (i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, c_ptr, c_ptr]),
}


class GsrLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises when the library is
    missing -- the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GsrLibraryError(
                f"{LIB_PATH} not found: build it with `python {os.path.join(_HERE, 'build.py')}` "
                "(nvcc, sm_100a).  There is no CPU / eager fallback for the GSR-Net path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        got = handle.gsr_abi_version()
        if got != ABI_VERSION:
            raise GsrLibraryError(f"libgsr_sm100 ABI {got} != binding ABI {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().gsr_last_error().decode(errors="replace")
        raise GsrLibraryError(f"{what or 'libgsr_sm100'} failed with code {rc}: {msg}")


def launch_count() -> int:
    return int(lib().gsr_launch_count())


def set_tuning(key: str, value: int) -> None:
    check(lib().gsr_set_tuning(key.encode(), int(value)), "gsr_set_tuning")
