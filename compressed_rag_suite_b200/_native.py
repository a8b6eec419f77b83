"""ctypes binding of libcrs.so (include/crs.h) — the only way Python reaches the GPU path.

There is no Python/NumPy implementation behind these calls: if the shared library
is missing this module raises at import, and if no sm_100 device is usable
``crs_index_create`` fails with CRS_ECUDA and the wrapper raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrs.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

CRS_F32, CRS_F16, CRS_BF16, CRS_I8, CRS_B1 = 0, 1, 2, 3, 4
CRS_COSINE, CRS_IP = 0, 1
CRS_OK, CRS_EINVAL, CRS_ECUDA, CRS_ENOMEM, CRS_ESTATE, CRS_EIO = 0, 1, 2, 3, 4, 5
CRS_PAD_ID = 0xFFFFFFFF

DTYPE_CODES = {"f32": CRS_F32, "f16": CRS_F16, "bf16": CRS_BF16, "i8": CRS_I8, "b1": CRS_B1}
METRIC_CODES = {"cosine": CRS_COSINE, "ip": CRS_IP}

# every symbol include/crs.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "crs_last_error": (C.c_char_p, []),
    "crs_version": (C.c_int, []),
    "crs_index_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int64]),
    "crs_index_destroy": (C.c_int, [_P]),
    "crs_index_set_stream": (C.c_int, [_P, _P]),
    "crs_index_add": (C.c_int, [_P, _P, C.c_int64, C.c_int]),
    "crs_index_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "crs_index_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "crs_index_search": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
    "crs_index_search_filtered": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P]),
    "crs_index_last_stats": (C.c_int, [_P, _P]),
    "crs_index_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "crs_index_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "crs_index_kernel_ms_history": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "crs_index_similarity_scale": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "crs_index_fetch_rows": (C.c_int, [_P, _P, C.c_int, _P]),
    "crs_index_score_rows": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    "crs_index_score_vectors": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    "crs_select_topk": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "crs_mmr": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_double, _P]),
    "crs_mmr_select": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P, _P, _P]),
    "crs_merge_topk": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "crs_merge_topk_strided": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P, _P]),
    "crs_exchange_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "crs_exchange_destroy": (C.c_int, [_P]),
    "crs_exchange_ipc_handle": (C.c_int, [_P, _P]),
    "crs_exchange_open_peers": (C.c_int, [_P, _P]),
    "crs_exchange_buffer": (C.c_int, [_P, C.POINTER(_P)]),
    "crs_exchange_set_peer_buffers": (C.c_int, [_P, C.POINTER(_P)]),
    "crs_exchange_status": (C.c_int, [_P, _P, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]),
    "crs_index_search_sharded": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
    "crs_index_search_push": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_float, _P]),
    "crs_index_map_ids": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_uint32]),
    "crs_exchange_merge": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "crs_index_codes_handle": (C.c_int, [_P, _P, C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)]),
    "crs_exchange_open_shards": (C.c_int, [_P, _P, _P, C.POINTER(C.c_uint32), C.POINTER(C.c_int64)]),
    "crs_exchange_set_shards": (C.c_int, [_P, _P, C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)]),
    "crs_exchange_fetch_rows": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "crs_exchange_score_rows": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, _P]),
    "crs_index_save": (C.c_int, [_P, C.c_char_p]),
    "crs_index_append": (C.c_int, [_P, C.c_char_p]),
    "crs_index_truncate": (C.c_int, [_P, C.c_int64]),
    "crs_index_load": (C.c_int, [C.POINTER(_P), C.c_char_p, C.c_int, C.c_uint32]),
}


class SearchStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int32), ("path", C.c_int32), ("grid", C.c_int32),
                ("list_len", C.c_int32), ("uncertified_total", C.c_int64), ("searches_total", C.c_int64),
                ("max_fast_error", C.c_float), ("reserved", C.c_int32)]


def build_library(force: bool = False) -> str:
    """Compile libcrs.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.run(["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)], check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load libcrs.so (building it if the .so is absent and nvcc is present)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            try:
                build_library()
            except Exception as e:          # no nvcc / build failure: there is no fallback
                raise RuntimeError(
                    f"libcrs.so is missing and could not be built ({e}); the CUDA library is required — "
                    "there is no CPU fallback") from e
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class CrsError(RuntimeError):
    pass


def check(status: int) -> None:
    """Non-zero status -> the exception the reference would raise after logging
    (rag/indexing.py:121-123,178-180): ValueError for bad arguments, RuntimeError otherwise."""
    if status == CRS_OK:
        return
    msg = (lib().crs_last_error() or b"").decode("utf-8", "replace")
    if status == CRS_EINVAL:
        raise ValueError(f"crs: {msg}")
    if status == CRS_ENOMEM:
        raise MemoryError(f"crs: {msg}")
    raise CrsError(f"crs (status {status}): {msg}")
