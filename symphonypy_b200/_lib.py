"""ctypes binding of the C-ABI CUDA library (`include/symphony_b200.h`).

There is no CPU fallback: if the library is missing the import of any product
module that needs it fails loudly, telling the user how to build it.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# SYM_B200_LIB: experiments only (an instrumented / tuning build made by `python -m symphonypy_b200.build --variant`)
LIB_PATH = os.environ.get("SYM_B200_LIB") or os.path.join(_HERE, "_C", "libsymphony_b200.so")

# symbol -> (restype, argtypes); mirrors include/symphony_b200.h line by line
_P = c_void_p
SIGNATURES = {
    "sym_version": (c_int32, []),
    "sym_last_error": (c_char_p, []),
    "sym_launch_count": (c_int64, []),
    "sym_compiled_arch": (c_int32, []),
    "sym_host_digest64": (c_uint64, [_P, c_uint64, c_int32]),
    "sym_host_copy": (None, [_P, _P, c_uint64, c_int32]),
    "sym_project_dense": (c_int32, [_P, c_int64, c_int64, _P, c_int32, _P, _P, _P, c_int32, c_int32, c_float, _P,
                                    c_int32, _P]),
    "sym_project_pack_bytes": (c_size_t, [c_int32, c_int32]),
    "sym_project_pack": (c_int32, [_P, c_int32, c_int32, c_int32, c_float, _P, _P]),
    "sym_project_dense_tc": (c_int32, [_P, c_int64, c_int64, c_int32, _P, _P, _P, c_int32, c_float, c_float, _P, c_int32,
                                      _P]),
    "sym_project_csr_bias": (c_int32, [_P, _P, _P, c_int32, c_int32, c_int32, c_float, _P, _P]),
    "sym_project_csr": (c_int32, [_P, _P, _P, c_int64, _P, _P, _P, _P, c_int32, c_int32, c_float, _P, _P, c_int32, _P]),
    "sym_project_csr_checked": (c_int32, [_P, _P, _P, c_int64, _P, _P, _P, _P, c_int32, c_int32, c_float, _P, _P, c_int32,
                                          _P, _P]),
    "sym_assign": (c_int32, [_P, c_int32, c_int64, c_int32, _P, c_int32, _P, _P, _P]),
    "sym_batch_order_workspace": (c_size_t, [c_int32]),
    "sym_batch_order": (c_int32, [_P, c_int64, c_int32, _P, _P, _P, _P, c_size_t, _P]),
    "sym_moe_stats": (c_int32, [_P, c_int32, _P, c_int64, c_int32, c_int32, _P, c_int32, _P, c_int32, c_int32, _P, _P,
                                _P, c_int32, _P, _P, _P]),
    "sym_moe_solve": (c_int32, [_P, _P, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P, _P]),
    "sym_moe_apply": (c_int32, [_P, c_int32, _P, c_int64, c_int32, c_int32, _P, c_int32, c_int32, _P, _P, _P, _P,
                                c_int32, _P, c_int32, _P]),
    "sym_knn_packed_bytes": (c_size_t, [c_int64, c_int32]),
    "sym_knn_fit": (c_int32, [_P, c_int32, c_int64, c_int32, _P, _P, _P]),
    "sym_kmeanspp_workspace": (c_size_t, [c_int64, c_int32]),
    "sym_kmeanspp": (c_int32, [_P, c_int32, c_int64, c_int32, c_int32, c_int32, c_int64, _P, _P, _P, c_size_t, _P]),
    "sym_nearest_centroid": (c_int32, [_P, c_int32, c_int64, c_int32, _P, c_int32, _P, _P]),
    "sym_knn_workspace": (c_size_t, [c_int64, c_int64, c_int32, c_int32, c_int32]),
    "sym_knn_method_supported": (c_int32, [c_int64, c_int64, c_int32, c_int32, c_int32]),
    "sym_knn": (c_int32, [_P, c_int32, c_int64, _P, c_int32, c_int64, _P, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P,
                          _P, _P, c_size_t, _P]),
    "sym_knn_exact_rows_max": (c_int32, []),
    "sym_knn_exact_rows_workspace": (c_size_t, [c_int32]),
    "sym_knn_exact_rows": (c_int32, [_P, c_int32, _P, c_int32, _P, c_int32, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P,
                                     c_size_t, _P]),
    "sym_vote": (c_int32, [_P, _P, c_int64, c_int32, _P, c_int64, _P, _P, _P]),
    "sym_cluster_moments": (c_int32, [_P, c_int32, c_int64, c_int32, _P, c_int32, _P, _P]),
    "sym_cluster_cov": (c_int32, [_P, c_int32, c_int64, c_int32, _P, c_int32, _P, _P, _P]),
    "sym_cell_confidence": (c_int32, [_P, c_int32, c_int64, c_int32, _P, c_int32, _P, _P, _P, _P]),
    "sym_group_sums": (c_int32, [_P, c_int32, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P]),
    "sym_group_cov": (c_int32, [_P, c_int32, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P, _P]),
    "sym_debug_tc_scores": (c_int32, [_P, c_int32, _P, c_int32, c_int32, _P, _P, _P, c_int32]),
}

_lib = None


class SymphonyLibraryError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SymphonyLibraryError(
            f"{LIB_PATH} is missing. symphonypy_b200 has no CPU fallback: build the CUDA library with "
            "`python -m symphonypy_b200.build` (needs nvcc with sm_100a support).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sym_last_error().decode("utf-8", "replace")
        raise SymphonyLibraryError(f"{what or 'symphony_b200'} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().sym_launch_count())
