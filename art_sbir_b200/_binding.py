"""ctypes binding of libsbir_b200.so (the C ABI declared in include/sbir_b200.h).

The library is loaded from art_sbir_b200/lib/ (built in-tree by `_build.build()`).  There is
no fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libsbir_b200.so"

SBIR_F32, SBIR_BF16 = 0, 1
SBIR_EUCLIDEAN, SBIR_COSINE = 0, 1
MAX_K = 116
ABI_VERSION = 2

# name -> (restype, argtypes); mirrors include/sbir_b200.h one to one.
_P = c_void_p
PROTOTYPES = {
    "sbir_abi_version": (c_int, []),
    "sbir_status_string": (c_char_p, [c_int]),
    "sbir_last_cuda_error": (c_int, []),
    "sbir_device_supported": (c_int, []),
    "sbir_l2_normalize": (c_int, [_P, _P, c_int64, c_int64, c_int, c_float, _P]),
    "sbir_row_sqnorm": (c_int, [_P, c_int64, c_int64, c_int, _P, _P]),
    "sbir_pairwise_distance": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, c_int, _P, _P]),
    "sbir_pairwise_distance_bwd": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, _P, _P, _P, _P]),
    "sbir_pairwise_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int]),
    "sbir_gallery_append": (c_int, [_P, c_int, c_int64, c_int64, _P, c_int, c_int64, c_int64, _P, c_int, _P]),
    "sbir_pairwise_topk": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int64, _P, _P, _P,
                                   _P, _P, _P, c_size_t, _P]),
    "sbir_positive_distance": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, c_int, _P, _P, _P]),
    "sbir_pairwise_topk_shard": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int64, _P, _P, _P,
                                         _P, _P, _P, _P, c_size_t, _P]),
    "sbir_topk_merge": (c_int, [_P, _P, c_int, c_int64, c_int64, c_int64, c_int, _P, _P, _P]),
    "sbir_retrieval_metrics": (c_int, [_P, c_int64, c_int, _P, _P]),
    "sbir_triplet_margin_loss": (c_int, [_P, _P, _P, c_int64, c_int64, c_float, c_int, _P, _P, _P, _P, _P, _P]),
    "sbir_batch_hard_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "sbir_batch_hard_triplet_loss": (c_int, [_P, _P, _P, c_int64, c_int64, c_float, c_int, _P, _P, _P, _P, _P,
                                             _P, _P, _P, c_size_t, _P]),
    "sbir_retrieve_host": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "sbir_retrieve_host_shard": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, c_int64, _P, _P, _P, _P,
                                         _P, _P, _P]),
    "sbir_gather_rows_host": (c_int, [_P, c_int64, c_int64, _P, c_int64, _P, c_int]),
    "sbir_release_host_staging": (c_int, []),
    "sbir_profile_enable": (c_int, [c_int]),
    "sbir_profile_collect": (c_int, [_P, _P, _P]),
    "sbir_debug_set_option": (c_int, [c_char_p, c_int64]),
    "sbir_debug_diag_build": (c_int, []),
    "sbir_debug_plan": (c_int, [c_int64, c_int64, c_int64, c_int, c_int, c_int, _P]),
    "sbir_debug_k1_diag": (c_int, [_P, c_int]),
    "sbir_debug_dist_matrix_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "sbir_debug_dist_matrix": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, c_int, _P, _P, c_size_t, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load (once) and return the shared library with typed prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m art_sbir_b200._build` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sbir_abi_version() != ABI_VERSION:
        raise RuntimeError("libsbir_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        lib = load()
        msg = lib.sbir_status_string(status).decode()
        extra = f" (cudaError {lib.sbir_last_cuda_error()})" if status == 3 else ""
        raise RuntimeError(f"{what}: {msg}{extra}")


def set_debug_option(name: str, value: int = 0) -> None:
    """Process-wide tuning / test switch of the library (include/sbir_b200.h: sbir_debug_set_option)."""
    check(load().sbir_debug_set_option(name.encode(), int(value)), f"sbir_debug_set_option({name})")
