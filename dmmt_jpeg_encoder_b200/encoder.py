"""Host-side objects over the C ABI: contexts, plans, batches.

PyTorch is only the plumbing here (device memory, streams); every byte of the JPEG is produced by
the sm_100a kernels behind include/dmmt_cuda.h.  Nothing in this module can run without the CUDA
library and a GPU: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi as F


@dataclass(frozen=True)
class Options:
    """JpegTransformationOptions (reference src/image/writer/jpeg.rs:25-29) in C-ABI form."""
    subsampling: int = F.P420
    bits_per_channel: int = 8
    qtable_preset: int = 0

    def c(self) -> F.Options:
        return F.Options(self.subsampling, self.bits_per_channel, self.qtable_preset)


def _fmt_of(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.uint8:
        return F.FMT_U8
    if dtype == np.uint16:
        return F.FMT_U16
    if dtype == np.float32:
        return F.FMT_F32_NORM
    raise TypeError(f"pixels must be uint8, uint16 or float32 (normalised), got {dtype}")


class Context:
    """dmmt_ctx: one device + one stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        L = F.lib()
        if stream is None:
            F.check(L.dmmt_ctx_create(device, C.byref(self._h)), "dmmt_ctx_create")
        else:
            F.check(L.dmmt_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(self._h)),
                    "dmmt_ctx_create_on_stream")
        self.device = device

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        return F.lib().dmmt_ctx_stream(self._h) or 0

    def synchronize(self):
        F.check(F.lib().dmmt_ctx_synchronize(self._h), "dmmt_ctx_synchronize")

    def debug_stuff(self, scan: bytes, misalign: int = 0) -> bytes:
        """dmmt_debug_stuff: K4 alone on an unstuffed scan -> stuffed bytes + EOI (test hook)."""
        src = np.frombuffer(scan, np.uint8) if len(scan) else np.zeros(1, np.uint8)
        dst = np.empty(2 * len(scan) + 64, np.uint8)
        got = C.c_size_t(0)
        F.check(F.lib().dmmt_debug_stuff(self._h, src.ctypes.data_as(C.c_void_p), len(scan), misalign,
                                         dst.ctypes.data_as(C.c_void_p), dst.size, C.byref(got)), "dmmt_debug_stuff")
        return dst[:got.value].tobytes()

    def encode(self, pixels: np.ndarray, max_value: int = 255, options: Options = Options()) -> bytes:
        """dmmt_encode on host pixels [H, W, 3] (u8 / u16 / f32 normalised)."""
        px = np.ascontiguousarray(pixels)
        h, w, c = px.shape
        if c != 3:
            raise ValueError("pixels must be [H, W, 3]")
        if h > 65535 or w > 65535:
            raise F.DmmtError(F.E_SIZE, "dmmt_encode")
        im = F.Image(w, h, max_value, _fmt_of(px.dtype), px.ctypes.data, 0)
        return self._encode_image(im, options)

    def encode_device(self, d_ptr: int, width: int, height: int, fmt: int, max_value: int = 255,
                      options: Options = Options()) -> bytes:
        im = F.Image(width, height, max_value, fmt, d_ptr, 1)
        return self._encode_image(im, options)

    def _encode_image(self, im: F.Image, options: Options) -> bytes:
        out = F._U8P()
        n = C.c_size_t()
        o = options.c()
        F.check(F.lib().dmmt_encode(self._h, C.byref(im), C.byref(o), C.byref(out), C.byref(n)), "dmmt_encode")
        try:
            return C.string_at(out, n.value)
        finally:
            F.lib().dmmt_free(out)

    def encode_sharded(self, pixels: np.ndarray, n_shards: int, max_value: int = 255,
                       options: Options = Options(), contexts: list["Context"] | None = None) -> bytes:
        """dmmt_encode_sharded: MCU-row shards on `contexts` (default: n_shards times this context)."""
        px = np.ascontiguousarray(pixels)
        h, w, _ = px.shape
        ctxs = contexts or [self] * n_shards
        arr = (C.c_void_p * len(ctxs))(*[c.handle for c in ctxs])
        im = F.Image(w, h, max_value, _fmt_of(px.dtype), px.ctypes.data, 0)
        out = F._U8P()
        n = C.c_size_t()
        o = options.c()
        F.check(F.lib().dmmt_encode_sharded(arr, len(ctxs), C.byref(im), C.byref(o), C.byref(out), C.byref(n)),
                "dmmt_encode_sharded")
        try:
            return C.string_at(out, n.value)
        finally:
            F.lib().dmmt_free(out)

    def close(self):
        if self._h:
            F.lib().dmmt_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """dmmt_plan: fixed geometry, one launch chain over up to n images."""

    def __init__(self, ctx: Context, width: int, height: int, fmt: int = F.FMT_U8, max_value: int = 255,
                 options: Options = Options(), n_images: int = 1):
        self.ctx = ctx
        self._h = C.c_void_p()
        o = options.c()
        if width > 65535 or height > 65535:
            raise F.DmmtError(F.E_SIZE, "dmmt_plan_create")
        F.check(F.lib().dmmt_plan_create(ctx.handle, width, height, fmt, max_value, C.byref(o), n_images,
                                         C.byref(self._h)), "dmmt_plan_create")
        self.n = n_images
        self.width, self.height, self.fmt = width, height, fmt

    @property
    def handle(self):
        return self._h

    @property
    def pixel_bytes(self) -> int:
        return F.lib().dmmt_plan_pixel_bytes(self._h)

    @property
    def out_stride(self) -> int:
        return F.lib().dmmt_plan_out_stride(self._h)

    @property
    def stream_blocks(self) -> int:
        return F.lib().dmmt_plan_stream_blocks(self._h)

    def set_scan_capacity(self, bytes_per_image: int):
        F.check(F.lib().dmmt_plan_set_scan_capacity(self._h, bytes_per_image), "dmmt_plan_set_scan_capacity")

    def worst_case_scan_bytes(self) -> int:
        return F.lib().dmmt_plan_worst_case_scan_bytes(self._h)

    def set_profiling(self, on: bool):
        F.check(F.lib().dmmt_plan_set_profiling(self._h, int(on)), "dmmt_plan_set_profiling")

    def set_graph(self, on: bool):
        F.check(F.lib().dmmt_plan_set_graph(self._h, int(on)), "dmmt_plan_set_graph")

    def set_generic_path(self, on: bool):
        F.check(F.lib().dmmt_plan_set_generic_path(self._h, int(on)), "dmmt_plan_set_generic_path")

    def last_timings(self) -> dict:
        ms = (C.c_float * F.T_COUNT)()
        F.check(F.lib().dmmt_plan_last_timings(self._h, ms, F.T_COUNT), "dmmt_plan_last_timings")
        return dict(zip(F.T_NAMES, [float(v) for v in ms]))

    def last_launch_count(self) -> int:
        return F.lib().dmmt_plan_last_launch_count(self._h)

    def encode_device(self, d_pixels: int, n_images: int, d_out: int, d_lens: int = 0):
        F.check(F.lib().dmmt_plan_encode_device(self._h, C.c_void_p(d_pixels), n_images, C.c_void_p(d_out),
                                                C.c_void_p(d_lens) if d_lens else None),
                "dmmt_plan_encode_device")

    def status(self):
        F.check(F.lib().dmmt_plan_status(self._h), "dmmt_plan_status")

    def encode_host(self, pixels: np.ndarray) -> list[bytes]:
        """pixels: [n, H, W, 3] host array -> list of JPEG files."""
        px = np.ascontiguousarray(pixels)
        n = px.shape[0]
        assert px.nbytes == n * self.pixel_bytes, (px.shape, self.pixel_bytes)
        outs = (F._U8P * n)()
        lens = (C.c_size_t * n)()
        F.check(F.lib().dmmt_plan_encode_host(self._h, C.c_void_p(px.ctypes.data), n, outs, lens),
                "dmmt_plan_encode_host")
        res = []
        for i in range(n):
            res.append(C.string_at(outs[i], lens[i]))
            F.lib().dmmt_free(outs[i])
        return res

    def encode_host_into(self, h_pixels: int, n: int, h_out: int, out_cap: int, offsets: np.ndarray, lens: np.ndarray):
        """dmmt_plan_encode_host_into: host pixels (pointer, ideally pinned) -> packed files in the caller's host arena."""
        assert offsets.dtype == np.uint64 and lens.dtype == np.uint64 and len(offsets) >= n and len(lens) >= n
        F.check(F.lib().dmmt_plan_encode_host_into(self._h, C.c_void_p(h_pixels), n, C.c_void_p(h_out), out_cap,
                                                   offsets.ctypes.data_as(F._U64P), lens.ctypes.data_as(F._U64P)),
                "dmmt_plan_encode_host_into")

    def fetch(self, what: int, index: int = 0):
        L = F.lib()
        got = C.c_size_t()
        if what == F.FETCH_META:
            m = F.Meta()
            F.check(L.dmmt_plan_fetch(self._h, what, index, C.byref(m), C.sizeof(m), C.byref(got)), "dmmt_plan_fetch")
            return m
        if what == F.FETCH_TOKEN_COUNT:
            n = C.c_uint64()
            F.check(L.dmmt_plan_fetch(self._h, what, index, C.byref(n), 8, C.byref(got)), "dmmt_plan_fetch")
            return int(n.value)
        if what == F.FETCH_COEF:
            a = np.empty((self.stream_blocks, 64), np.int16)
        elif what == F.FETCH_HIST:
            a = np.empty((4, 256), np.uint32)
        elif what == F.FETCH_TABLES:
            a = np.empty((4, 2, 256), np.uint8)
        elif what == F.FETCH_SCAN:
            a = np.empty(self.out_stride, np.uint8)
        else:
            raise ValueError(what)
        F.check(L.dmmt_plan_fetch(self._h, what, index, C.c_void_p(a.ctypes.data), a.nbytes, C.byref(got)),
                "dmmt_plan_fetch")
        if what == F.FETCH_SCAN:
            return a[: got.value].copy()
        if what == F.FETCH_TABLES:
            # [4][sym 256 | len 256] is LenTables{sym[4][256], len[4][256]}: regroup per table
            raw = a.reshape(2, 4, 256)
            return raw[0], raw[1]
        return a

    def debug_dct(self, d_pixels: int, index: int = 0) -> np.ndarray:
        a = np.empty((self.stream_blocks, 64), np.float32)
        F.check(F.lib().dmmt_plan_debug_dct(self._h, C.c_void_p(d_pixels), index,
                                            a.ctypes.data_as(C.POINTER(C.c_float)), a.size), "dmmt_plan_debug_dct")
        return a

    def close(self):
        if self._h:
            F.lib().dmmt_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """dmmt_batch: pipelined sub-batches over `depth` streams (the throughput path)."""

    def __init__(self, ctx: Context, width: int, height: int, fmt: int = F.FMT_U8, max_value: int = 255,
                 options: Options = Options(), sub_batch: int = 32, depth: int = 3):
        self.ctx = ctx
        self._h = C.c_void_p()
        o = options.c()
        F.check(F.lib().dmmt_batch_create(ctx.handle, width, height, fmt, max_value, C.byref(o), sub_batch, depth,
                                          C.byref(self._h)), "dmmt_batch_create")
        self.width, self.height, self.fmt = width, height, fmt
        self.sub_batch, self.depth = sub_batch, depth

    def set_scan_capacity(self, bytes_per_image: int):
        F.check(F.lib().dmmt_batch_set_scan_capacity(self._h, bytes_per_image), "dmmt_batch_set_scan_capacity")

    def worst_case_scan_bytes(self) -> int:
        return F.lib().dmmt_batch_worst_case_scan_bytes(self._h)

    def set_profiling(self, on: bool):
        F.check(F.lib().dmmt_batch_set_profiling(self._h, int(on)), "dmmt_batch_set_profiling")

    def last_timings(self) -> dict:
        ms = (C.c_float * F.T_COUNT)()
        F.check(F.lib().dmmt_batch_last_timings(self._h, ms, F.T_COUNT), "dmmt_batch_last_timings")
        return dict(zip(F.T_NAMES, [float(v) for v in ms]))

    def last_launch_count(self) -> int:
        return F.lib().dmmt_batch_last_launch_count(self._h)

    def uses_fused_path(self) -> bool:
        return bool(F.lib().dmmt_batch_uses_fused_path(self._h))

    def encode_device(self, d_pixels: int, n: int, d_dense: int, dense_cap: int, d_offsets: int, d_lens: int):
        """Asynchronous; all pointers are device pointers (offsets: n + 1 u64, lens: n u64)."""
        F.check(F.lib().dmmt_batch_encode_device(self._h, C.c_void_p(d_pixels), n, C.c_void_p(d_dense), dense_cap,
                                                 C.c_void_p(d_offsets), C.c_void_p(d_lens)),
                "dmmt_batch_encode_device")

    def status(self):
        F.check(F.lib().dmmt_batch_status(self._h), "dmmt_batch_status")

    def encode_host(self, h_pixels: int, n: int, h_out: int, out_cap: int, offsets: np.ndarray, lens: np.ndarray):
        """Synchronous: host pixels (pointer) -> packed files in the host arena (pointer)."""
        assert offsets.dtype == np.uint64 and lens.dtype == np.uint64 and len(offsets) >= n and len(lens) >= n
        F.check(F.lib().dmmt_batch_encode_host(self._h, C.c_void_p(h_pixels), n, C.c_void_p(h_out), out_cap,
                                               offsets.ctypes.data_as(F._U64P), lens.ctypes.data_as(F._U64P)),
                "dmmt_batch_encode_host")

    def encode(self, pixels: np.ndarray) -> list[bytes]:
        """Convenience: [n, H, W, 3] host array -> list of files."""
        px = np.ascontiguousarray(pixels)
        n = px.shape[0]
        cap = max(1 << 20, px.nbytes // 2 + 4096 * n)
        while True:
            out = np.empty(cap, np.uint8)
            offs = np.zeros(n, np.uint64)
            lens = np.zeros(n, np.uint64)
            try:
                self.encode_host(px.ctypes.data, n, out.ctypes.data, cap, offs, lens)
            except F.DmmtError as e:
                if e.code == F.E_WRITE and cap < 4 * px.nbytes + (1 << 22):
                    cap *= 4
                    continue
                if e.code == F.E_OVERFLOW:
                    self.set_scan_capacity(self.worst_case_scan_bytes())
                    continue
                raise
            return [out[int(o): int(o) + int(l)].tobytes() for o, l in zip(offs, lens)]

    def close(self):
        if self._h:
            F.lib().dmmt_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """dmmt_host_alloc'd page-locked host memory exposed as a numpy u8 array."""

    def __init__(self, nbytes: int):
        self._p = C.c_void_p()
        F.check(F.lib().dmmt_host_alloc(nbytes, C.byref(self._p)), "dmmt_host_alloc")
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(C.cast(self._p, F._U8P), shape=(max(nbytes, 1),))[:nbytes]

    @property
    def ptr(self) -> int:
        return self._p.value

    def close(self):
        if self._p:
            self.array = None
            F.lib().dmmt_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
