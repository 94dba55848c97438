"""ctypes binding to libcadence_dense.so (C ABI: include/cadence_dense.h).

This is the reference-side stub a maintainer adds: every function declared in
the header is bound here with explicit argtypes; errors become :class:`DenseEngineError`.
There is deliberately no fallback: if the shared library is missing, or it reports no CUDA
device, calls raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libcadence_dense.so"

CDR_OK = 0
CDR_ERR_INVALID = -1
CDR_ERR_CUDA = -2
CDR_ERR_OOM = -3
CDR_ERR_UNSORTED_IDS = -4
CDR_ERR_STATE = -5
CDR_ERR_UNSUPPORTED = -6
CDR_ERR_NO_DEVICE = -7

CDR_STORE_FP32 = 1
CDR_STORE_BF16 = 2
CDR_MAX_K = 248
CDR_RRF_MAX_ITEMS = 1024
CDR_PEER_MAX_RANKS = 16
CDR_PEER_HANDLE_BYTES = 64

_CODE_NAMES = {
    CDR_ERR_INVALID: "CDR_ERR_INVALID", CDR_ERR_CUDA: "CDR_ERR_CUDA", CDR_ERR_OOM: "CDR_ERR_OOM",
    CDR_ERR_UNSORTED_IDS: "CDR_ERR_UNSORTED_IDS", CDR_ERR_STATE: "CDR_ERR_STATE",
    CDR_ERR_UNSUPPORTED: "CDR_ERR_UNSUPPORTED", CDR_ERR_NO_DEVICE: "CDR_ERR_NO_DEVICE",
}


class DenseEngineError(RuntimeError):
    """Raised for every non-zero status of the C ABI (and when the library cannot be loaded).

    The retrieval facade treats it like the reference treats EmbeddingClientError
    (app/retrieve.py:426-432): the dense lane is disabled for the request and the message is
    surfaced as ``dense_error``."""

    def __init__(self, message: str, code: int = CDR_ERR_INVALID):
        super().__init__(message)
        self.code = code


def library_path() -> str:
    return os.environ.get("CADENCE_DENSE_LIB", os.path.join(_HERE, _LIB_NAME))


_lib: Optional[ctypes.CDLL] = None

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_u32 = ctypes.c_uint32
_u64 = ctypes.c_uint64

# name -> (restype, argtypes); one entry per function declared in include/cadence_dense.h
SIGNATURES = {
    "cdr_abi_version": (_i32, []),
    "cdr_last_error": (ctypes.c_char_p, []),
    "cdr_device_count": (_i32, [ctypes.POINTER(_i32)]),
    "cdr_store_create": (_i32, [ctypes.POINTER(_vp), _i32, _i64, _i32, _u32]),
    "cdr_store_destroy": (_i32, [_vp]),
    "cdr_store_append": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "cdr_store_update_embeddings": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "cdr_store_append_synthetic": (_i32, [_vp, _u64, _i64, _i64, _i64, _i32, _i64, _i64, _vp]),
    "cdr_store_finalize": (_i32, [_vp, _vp]),
    "cdr_store_info": (_i32, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(_i32), ctypes.POINTER(_u32),
                              ctypes.POINTER(_i64), ctypes.POINTER(_i32)]),
    "cdr_store_read_rows": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cdr_store_copy_rows_device": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "cdr_store_read_valid": (_i32, [_vp, _i64, _i64, _vp]),
    "cdr_synth_rows": (_i32, [_vp, _u64, _i64, _i64, _i32, _vp]),
    "cdr_filter_build": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _i64, _i32, _u64, _vp,
                                ctypes.POINTER(_i64), _vp]),
    "cdr_search_exact_f32": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_search_exact_f32_host": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_search_exact_f32_shared": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_search_exact_f32_shared_host": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_search_scan_bf16": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_search_scan_bf16_host": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_search_batch_bf16": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_search_batch_bf16_host": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_topk_merge": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "cdr_peer_group_create": (_i32, [ctypes.POINTER(_vp), _i32, _i32, _i32, _i32, _i32, _vp]),
    "cdr_peer_group_connect": (_i32, [_vp, _vp]),
    "cdr_peer_group_destroy": (_i32, [_vp]),
    "cdr_peer_exchange_merge": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "cdr_search_sharded": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_rrf_merge": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_rrf_merge_host": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cdr_tech_index_create": (_i32, [ctypes.POINTER(_vp), _vp, _vp, _i32, _vp, _vp]),
    "cdr_tech_index_destroy": (_i32, [_vp]),
    "cdr_tech_lane_host": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _i64, _i32, _i64, _i32, _u64, _i32,
                                  _vp, _vp, _vp]),
    "cdr_hybrid_retrieve_host": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cdr_hybrid_retrieve_groups_host": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp,
                                               _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cdr_kernel_launch_count": (_i64, []),
    "cdr_prof_enable": (_i32, [_i32]),
    "cdr_prof_read": (_i32, [_i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64)]),
    "cdr_prof_read_launches": (_i32, [_i32, _vp, _i64, ctypes.POINTER(_i64)]),
}


CDR_DENSE_LANE_EXACT_F32 = 0
CDR_DENSE_LANE_BATCH_BF16 = 1
CDR_DENSE_LANE_SCAN_BF16 = 2
CDR_DENSE_LANE_EXACT_F32_SHARED = 3     # cdr_search_sharded only


class FilterSpec(ctypes.Structure):
    """struct cdr_filter_spec (include/cadence_dense.h)."""
    _fields_ = [("call_slot_bitmap_host", _vp), ("n_call_slots", _i64), ("has_date_from", _i32),
                ("has_date_to", _i32), ("date_from_us", _i64), ("date_to_us", _i64), ("has_tag_filter", _i32),
                ("dense_lane", _i32), ("tag_any", _u64)]


def lib() -> ctypes.CDLL:
    """Load (once) and return the bound library; raises DenseEngineError if it is not built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise DenseEngineError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(sh cadence_rag_b200/csrc/build.sh). There is no CPU fallback.", CDR_ERR_STATE)
        try:
            L = ctypes.CDLL(path)
        except OSError as exc:
            raise DenseEngineError(f"cannot load {path}: {exc}", CDR_ERR_STATE) from exc
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    msg = lib().cdr_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str = "") -> None:
    if status != CDR_OK:
        name = _CODE_NAMES.get(status, str(status))
        raise DenseEngineError(f"{what + ': ' if what else ''}{name}: {last_error()}", status)


def abi_version() -> int:
    return int(lib().cdr_abi_version())


def kernel_launch_count() -> int:
    return int(lib().cdr_kernel_launch_count())


def device_count() -> int:
    n = _i32(0)
    check(lib().cdr_device_count(ctypes.byref(n)), "cdr_device_count")
    return int(n.value)


def require_device() -> None:
    """Fail loudly when the CUDA path cannot run (no silent eager/CPU path exists)."""
    device_count()


def stream_ptr(stream=None) -> int:
    """cudaStream_t of a torch stream (default: torch's current stream) as an integer."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return int(stream.cuda_stream)


def current_stream_ptr(device_index: int) -> int:
    """cudaStream_t of torch's current stream on a device, without building a torch.cuda.Stream object (this sits on the
    path of every search call, ahead of the first kernel launch)."""
    import torch
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return int(raw(device_index))
    return int(torch.cuda.current_stream(device_index).cuda_stream)


def ptr(t) -> Optional[int]:
    """data pointer of a torch tensor / numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return int(t.data_ptr())
    return int(t.ctypes.data)
