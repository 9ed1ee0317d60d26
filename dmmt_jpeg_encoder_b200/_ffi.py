"""ctypes binding of libdmmt_cuda.so (include/dmmt_cuda.h).

There is no CPU fallback: if the library is missing it is built with nvcc (build.py); if it
cannot be loaded, or no CUDA device is usable, the calls raise DmmtError.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

# error codes (include/dmmt_cuda.h)
OK, E_INVALID, E_NODEVICE, E_CUDA, E_NCCL, E_NOMEM, E_OVERFLOW, E_SYMBOL, E_RANGE, E_WRITE, E_SIZE = (
    0, -1, -2, -3, -4, -5, -6, -7, -8, -9, -10)
P444, P422, P420 = 0, 1, 2
FMT_F32_NORM, FMT_U8, FMT_U16 = 0, 1, 2
FETCH_COEF, FETCH_HIST, FETCH_TABLES, FETCH_SCAN, FETCH_META, FETCH_TOKEN_COUNT = 0, 1, 2, 3, 4, 5
T_NAMES = ("k1_transform", "k2_histogram", "k2b_tables", "k3_pack", "k4_stuff", "k5_compact", "total")
T_COUNT = 7


class DmmtError(RuntimeError):
    def __init__(self, code: int, where: str = ""):
        self.code = code
        msg = lib().dmmt_strerror(code).decode() if _lib is not None else str(code)
        detail = ""
        if code == E_CUDA and _lib is not None:
            detail = ": " + _lib.dmmt_last_cuda_error().decode()
        super().__init__(f"{where + ': ' if where else ''}{msg} ({code}){detail}")


class Image(C.Structure):
    _fields_ = [("width", C.c_uint16), ("height", C.c_uint16), ("max_value", C.c_uint16),
                ("fmt", C.c_int), ("pixels", C.c_void_p), ("pixels_on_device", C.c_int)]


class Options(C.Structure):
    _fields_ = [("subsampling", C.c_uint8), ("bits_per_channel", C.c_uint8), ("qtable_preset", C.c_uint8)]


class Meta(C.Structure):
    _fields_ = [("scan_bits", C.c_uint64), ("out_len", C.c_uint64), ("header_len", C.c_uint32),
                ("n_symbols", C.c_uint32 * 4), ("error", C.c_int32), ("n_stream_blocks", C.c_uint32),
                ("reserved", C.c_uint32)]


_VP = C.c_void_p
_PVP = C.POINTER(C.c_void_p)
_U8P = C.POINTER(C.c_uint8)
_U64P = C.POINTER(C.c_uint64)

# name -> (restype, argtypes): every symbol include/dmmt_cuda.h declares
SIGNATURES = {
    "dmmt_device_count": (C.c_int, []),
    "dmmt_ctx_create": (C.c_int, [C.c_int, _PVP]),
    "dmmt_ctx_create_on_stream": (C.c_int, [C.c_int, _VP, _PVP]),
    "dmmt_ctx_destroy": (None, [_VP]),
    "dmmt_ctx_synchronize": (C.c_int, [_VP]),
    "dmmt_ctx_stream": (_VP, [_VP]),
    "dmmt_encode": (C.c_int, [_VP, C.POINTER(Image), C.POINTER(Options), C.POINTER(_U8P), C.POINTER(C.c_size_t)]),
    "dmmt_encode_batch": (C.c_int, [_PVP, C.c_int, C.POINTER(Image), C.c_int, C.POINTER(Options),
                                    C.POINTER(_U8P), C.POINTER(C.c_size_t)]),
    "dmmt_encode_sharded": (C.c_int, [_PVP, C.c_int, C.POINTER(Image), C.POINTER(Options),
                                      C.POINTER(_U8P), C.POINTER(C.c_size_t)]),
    "dmmt_free": (None, [_VP]),
    "dmmt_strerror": (C.c_char_p, [C.c_int]),
    "dmmt_last_cuda_error": (C.c_char_p, []),
    "dmmt_ppm_parse": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16),
                                 C.POINTER(C.c_uint16), C.POINTER(C.POINTER(C.c_uint16)), C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_int)]),
    "dmmt_ppm_strerror": (C.c_char_p, [C.c_int, C.c_int, C.c_char_p, C.c_size_t]),
    "dmmt_plan_create": (C.c_int, [_VP, C.c_uint16, C.c_uint16, C.c_int, C.c_uint16, C.POINTER(Options), C.c_int, _PVP]),
    "dmmt_plan_destroy": (None, [_VP]),
    "dmmt_plan_pixel_bytes": (C.c_size_t, [_VP]),
    "dmmt_plan_out_stride": (C.c_size_t, [_VP]),
    "dmmt_plan_set_scan_capacity": (C.c_int, [_VP, C.c_size_t]),
    "dmmt_plan_worst_case_scan_bytes": (C.c_size_t, [_VP]),
    "dmmt_plan_encode_device": (C.c_int, [_VP, _VP, C.c_int, _VP, _VP]),
    "dmmt_plan_status": (C.c_int, [_VP]),
    "dmmt_plan_set_graph": (C.c_int, [_VP, C.c_int]),
    "dmmt_plan_encode_host": (C.c_int, [_VP, _VP, C.c_int, C.POINTER(_U8P), C.POINTER(C.c_size_t)]),
    "dmmt_plan_encode_host_into": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_uint64, _U64P, _U64P]),
    "dmmt_batch_create": (C.c_int, [_VP, C.c_uint16, C.c_uint16, C.c_int, C.c_uint16, C.POINTER(Options),
                                    C.c_int, C.c_int, _PVP]),
    "dmmt_batch_destroy": (None, [_VP]),
    "dmmt_batch_encode_device": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_uint64, _VP, _VP]),
    "dmmt_batch_encode_host": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_uint64, _U64P, _U64P]),
    "dmmt_batch_status": (C.c_int, [_VP]),
    "dmmt_batch_last_launch_count": (C.c_int, [_VP]),
    "dmmt_batch_set_profiling": (C.c_int, [_VP, C.c_int]),
    "dmmt_batch_last_timings": (C.c_int, [_VP, C.POINTER(C.c_float), C.c_int]),
    "dmmt_batch_set_scan_capacity": (C.c_int, [_VP, C.c_size_t]),
    "dmmt_batch_worst_case_scan_bytes": (C.c_size_t, [_VP]),
    "dmmt_host_alloc": (C.c_int, [C.c_size_t, _PVP]),
    "dmmt_host_free": (None, [_VP]),
    "dmmt_plan_set_profiling": (C.c_int, [_VP, C.c_int]),
    "dmmt_plan_set_generic_path": (C.c_int, [_VP, C.c_int]),
    "dmmt_plan_uses_fused_path": (C.c_int, [_VP]),
    "dmmt_batch_uses_fused_path": (C.c_int, [_VP]),
    "dmmt_plan_last_timings": (C.c_int, [_VP, C.POINTER(C.c_float), C.c_int]),
    "dmmt_plan_last_launch_count": (C.c_int, [_VP]),
    "dmmt_plan_fetch": (C.c_int, [_VP, C.c_int, C.c_int, _VP, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dmmt_plan_debug_dct": (C.c_int, [_VP, _VP, C.c_int, C.POINTER(C.c_float), C.c_size_t]),
    "dmmt_encode_sharded_last_ms": (C.c_double, []),
    "dmmt_mailbox_bytes": (C.c_size_t, [C.c_int]),
    "dmmt_shard_launch_post": (C.c_int, [_VP, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_ulonglong, _VP, C.c_int]),
    "dmmt_shard_launch_collect": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_ulonglong, C.c_int, C.c_int, _VP]),
    "dmmt_debug_stuff": (C.c_int, [_VP, _VP, C.c_size_t, C.c_int, _VP, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dmmt_plan_stream_blocks": (C.c_size_t, [_VP]),
    "dmmt_shard_create": (C.c_int, [_VP, C.c_uint16, C.c_uint16, C.c_int, C.c_uint16, C.POINTER(Options),
                                    C.c_int, C.c_int, _PVP]),
    "dmmt_shard_destroy": (None, [_VP]),
    "dmmt_shard_mcu_rows_total": (C.c_int, [C.c_uint16, C.POINTER(Options)]),
    "dmmt_shard_pixel_bytes": (C.c_size_t, [_VP]),
    "dmmt_shard_pixel_offset": (C.c_size_t, [_VP]),
    "dmmt_shard_transform": (C.c_int, [_VP, _VP, C.POINTER(C.c_int16)]),
    "dmmt_shard_histogram": (C.c_int, [_VP, C.POINTER(C.c_int16), _U64P]),
    "dmmt_shard_tables": (C.c_int, [_VP, _U64P, _U64P]),
    "dmmt_shard_pack": (C.c_int, [_VP, C.c_uint64, C.c_int, _U8P, C.POINTER(C.c_int)]),
    "dmmt_shard_stuff": (C.c_int, [_VP, C.c_uint8, C.c_int, C.c_int, C.c_int, C.POINTER(_VP), _U64P]),
    "dmmt_shard_launch_count": (C.c_int, [_VP]),
    "dmmt_shard_launch_transform": (C.c_int, [_VP, _VP, _VP]),
    "dmmt_shard_launch_histogram": (C.c_int, [_VP, _VP, _VP]),
    "dmmt_shard_launch_tables": (C.c_int, [_VP, _VP, _VP]),
    "dmmt_shard_launch_pack": (C.c_int, [_VP, _VP, C.c_int, _VP]),
    "dmmt_shard_launch_stuff": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, C.POINTER(_VP), _VP]),
    "dmmt_shard_status": (C.c_int, [_VP]),
    "dmmt_shard_launch_error": (C.c_int, [_VP, _VP]),
    "dmmt_shard_set_scan_capacity": (C.c_int, [_VP, C.c_size_t]),
    "dmmt_shard_worst_case_scan_bytes": (C.c_size_t, [_VP]),
    "dmmt_shard_out_stride": (C.c_size_t, [_VP]),
    "dmmt_device_alloc": (C.c_int, [_VP, C.c_size_t, C.POINTER(_VP)]),
    "dmmt_device_free": (C.c_int, [_VP, _VP]),
    "dmmt_peer_export": (C.c_int, [_VP, _VP, C.c_char_p]),
    "dmmt_peer_open": (C.c_int, [_VP, C.c_char_p, C.POINTER(_VP)]),
    "dmmt_peer_close": (C.c_int, [_VP, _VP]),
    "dmmt_shard_launch_count_bytes": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, _VP]),
    "dmmt_shard_launch_stuff_into": (C.c_int, [_VP, _VP, C.c_int, C.c_int, _VP, C.c_size_t, _VP, _VP]),
}

_lib = None


def library_path() -> str:
    return _build.SO


def lib():
    """Loads (building first if needed) the CUDA library.  Raises if that is impossible."""
    global _lib
    if _lib is None:
        path = os.environ.get("DMMT_CUDA_LIB") or _build.SO  # override: A/B builds of the same ABI
        if path == _build.SO and not os.path.exists(path):
            _build.build()
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError = header / library out of step
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, where: str = "") -> None:
    if rc != OK:
        raise DmmtError(rc, where)
