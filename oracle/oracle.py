"""ctypes binding of the CPU oracle (oracle/dmmt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

FMT_F32_NORM, FMT_U8, FMT_U16 = 0, 1, 2
P444, P422, P420 = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dmmt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Result(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("padded_width", C.c_int),
        ("padded_height", C.c_int), ("hr", C.c_int), ("vr", C.c_int),
        ("y_blocks", C.c_size_t), ("c_blocks", C.c_size_t), ("n_mcus", C.c_size_t),
        ("n_stream_blocks", C.c_size_t),
        ("dct_y", C.POINTER(C.c_float)), ("dct_cb", C.POINTER(C.c_float)),
        ("dct_cr", C.POINTER(C.c_float)), ("stream", C.POINTER(C.c_int16)),
        ("hist", (C.c_uint64 * 256) * 4), ("table_n", C.c_int * 4),
        ("table_sym", (C.c_uint8 * 256) * 4), ("table_len", (C.c_int * 256) * 4),
        ("scan_bits", C.c_uint64), ("scan_bytes_unstuffed", C.c_size_t),
        ("scan_bytes_stuffed", C.c_size_t), ("header_bytes", C.c_size_t),
        ("t_transform_s", C.c_double), ("t_encode_s", C.c_double),
    ]


class _BitWriter(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t), ("cap", C.c_size_t),
                ("buffer", C.c_uint8), ("used", C.c_uint8), ("init_val", C.c_uint8),
                ("stuff", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_normalize.restype = C.c_float
        L.orc_normalize.argtypes = [C.c_uint16, C.c_uint16]
        L.orc_rgb_to_ycbcr.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_fast_arai.argtypes = [C.POINTER(C.c_float), C.c_size_t]
        L.orc_dct8x8.argtypes = [C.POINTER(C.c_float)]
        L.orc_quantize.restype = C.c_int16
        L.orc_quantize.argtypes = [C.c_float, C.c_uint8]
        L.orc_qtable.restype = C.POINTER(C.c_uint8)
        L.orc_qtable.argtypes = [C.c_int, C.c_int]
        L.orc_zigzag.restype = C.POINTER(C.c_uint8)
        L.orc_categorize.argtypes = [C.c_int16, C.POINTER(C.c_uint16)]
        L.orc_rle_tokens.argtypes = [C.POINTER(C.c_int16), C.c_int, C.POINTER(C.c_uint8),
                                     C.POINTER(C.c_int16)]
        L.orc_padded_dims.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]
        L.orc_quadfold_index.restype = C.c_size_t
        L.orc_quadfold_index.argtypes = [C.c_size_t, C.c_size_t]
        FP = C.POINTER(C.c_float)
        L.orc_pad_image.restype = C.c_size_t
        L.orc_pad_image.argtypes = [FP, C.c_int, C.c_int, C.c_int, C.c_int, FP, C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]
        L.orc_subsample_value.argtypes = [FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, FP]
        L.orc_subsample_retile.restype = C.c_size_t
        L.orc_subsample_retile.argtypes = [FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, FP]
        L.orc_count_block.argtypes = [C.c_int16, C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.c_int,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_package_merge.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_int,
                                        C.POINTER(C.c_int)]
        L.orc_build_table.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_uint8), C.POINTER(C.c_int)]
        L.orc_canonical_codes.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.c_int,
                                          C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)]
        L.orc_bw_init.argtypes = [C.POINTER(_BitWriter), C.c_int, C.c_int]
        L.orc_bw_write_bits.argtypes = [C.POINTER(_BitWriter), C.POINTER(C.c_uint8), C.c_size_t]
        L.orc_bw_flush.argtypes = [C.POINTER(_BitWriter)]
        L.orc_bw_free.argtypes = [C.POINTER(_BitWriter)]
        L.orc_stuff_bytes.restype = C.c_size_t
        L.orc_stuff_bytes.argtypes = [C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_uint8)]
        L.orc_parse_ppm.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.POINTER(C.c_uint16))]
        L.orc_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_uint8)),
                                 C.POINTER(C.c_size_t), C.POINTER(_Result), C.c_int]
        L.orc_result_free.argtypes = [C.POINTER(_Result)]
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------- primitives
def normalize(v: int, mx: int) -> np.float32:
    return np.float32(lib().orc_normalize(v, mx))


def rgb_to_ycbcr(r, g, b):
    a = (C.c_float * 3)(r, g, b)
    o = (C.c_float * 3)()
    lib().orc_rgb_to_ycbcr(a, o)
    return np.array(o[:], dtype=np.float32)


def fast_arai(vec8) -> np.ndarray:
    a = np.ascontiguousarray(vec8, dtype=np.float32).copy()
    lib().orc_fast_arai(a.ctypes.data_as(C.POINTER(C.c_float)), 1)
    return a


def dct8x8(block64) -> np.ndarray:
    a = np.ascontiguousarray(block64, dtype=np.float32).reshape(64).copy()
    lib().orc_dct8x8(a.ctypes.data_as(C.POINTER(C.c_float)))
    return a


def quantize(d: float, q: int) -> int:
    return int(lib().orc_quantize(C.c_float(d), q))


def qtable(preset: int, chroma: bool) -> np.ndarray:
    p = lib().orc_qtable(preset, int(chroma))
    return np.array(p[:64], dtype=np.uint8)


def zigzag() -> np.ndarray:
    return np.array(lib().orc_zigzag()[:64], dtype=np.uint8)


def categorize(v: int):
    """-> (category, left-aligned u16 pattern); category -1 where the reference panics."""
    pat = C.c_uint16(0)
    cat = lib().orc_categorize(v, C.byref(pat))
    return cat, pat.value


def rle_tokens(seq):
    a = np.ascontiguousarray(seq, dtype=np.int16)
    z = (C.c_uint8 * (len(a) + 8))()
    v = (C.c_int16 * (len(a) + 8))()
    n = lib().orc_rle_tokens(a.ctypes.data_as(C.POINTER(C.c_int16)), len(a), z, v)
    return [(int(z[i]), int(v[i])) for i in range(n)]


def padded_dims(w, h, preset):
    pw, ph = C.c_int(), C.c_int()
    lib().orc_padded_dims(w, h, preset, C.byref(pw), C.byref(ph))
    return pw.value, ph.value


def quadfold_index(i, line_length):
    return int(lib().orc_quadfold_index(i, line_length))


def pad_image(rgb, nearest_w: int, nearest_h: int):
    """padder.rs:12-42 on normalised f32 dots [H, W, 3] -> padded [pH, pW, 3]."""
    a = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, _ = a.shape
    pw, ph = C.c_int(), C.c_int()
    FP = C.POINTER(C.c_float)
    n = lib().orc_pad_image(a.ctypes.data_as(FP), w, h, nearest_w, nearest_h, None, C.byref(pw), C.byref(ph))
    out = np.full((ph.value, pw.value, 3), np.nan, np.float32)
    assert n == pw.value * ph.value
    lib().orc_pad_image(a.ctypes.data_as(FP), w, h, nearest_w, nearest_h, out.ctypes.data_as(FP), None, None)
    return out


def subsample_value(plane, hr: int, vr: int, average: bool, sx: int, sy: int):
    """subsampling_iter().nth(sy).nth(sx) of subsampling.rs:169-229; None past the edge."""
    a = np.ascontiguousarray(plane, dtype=np.float32)
    h, w = a.shape
    v = C.c_float()
    rc = lib().orc_subsample_value(a.ctypes.data_as(C.POINTER(C.c_float)), w, h, hr, vr, int(average), sx, sy,
                                   C.byref(v))
    return None if rc < 0 else np.float32(v.value)


def subsample_retile(plane, hr: int, vr: int, average: bool, square: int) -> np.ndarray:
    """Subsampler::subsample_to_square_structure(square) (subsampling.rs:136-140)."""
    a = np.ascontiguousarray(plane, dtype=np.float32)
    h, w = a.shape
    out = np.full((w // hr) * (h // vr), np.nan, np.float32)
    FP = C.POINTER(C.c_float)
    n = lib().orc_subsample_retile(a.ctypes.data_as(FP), w, h, hr, vr, int(average), square, out.ctypes.data_as(FP))
    assert n == out.size
    return out


def count_blocks(blocks):
    """HuffmanCount::from_iter (symbol_counting.rs:55-74) over [(dc_value, [(zeros, value), ...]), ...]
    -> (dc_hist[16], ac_hist[256]) as the oracle's own counting loop produces them."""
    dc = np.zeros(16, np.uint64)
    ac = np.zeros(256, np.uint64)
    U64 = C.POINTER(C.c_uint64)
    for d, toks in blocks:
        z = (C.c_uint8 * max(1, len(toks)))(*[t[0] for t in toks])
        v = (C.c_int16 * max(1, len(toks)))(*[t[1] for t in toks])
        if lib().orc_count_block(d, z, v, len(toks), dc.ctypes.data_as(U64), ac.ctypes.data_as(U64)) < 0:
            raise ValueError("categorize panics")
    return dc, ac


def package_merge(sorted_freqs, limit):
    f = np.ascontiguousarray(sorted_freqs, dtype=np.uint64)
    out = (C.c_int * max(1, len(f)))()
    rc = lib().orc_package_merge(f.ctypes.data_as(C.POINTER(C.c_uint64)), len(f), limit, out)
    if rc < 0:
        raise ValueError(f"package_merge rc={rc}")
    return list(out[: len(f)])


def build_table(hist, limit=15, plus_one=True):
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    sym = (C.c_uint8 * 256)()
    ln = (C.c_int * 256)()
    n = lib().orc_build_table(h.ctypes.data_as(C.POINTER(C.c_uint64)), len(h), limit,
                              int(plus_one), sym, ln)
    if n < 0:
        raise ValueError(f"build_table rc={n}")
    return list(sym[:n]), list(ln[:n])


def canonical_codes(symbols, lengths):
    n = len(symbols)
    s = (C.c_uint8 * max(1, n))(*symbols)
    ln = (C.c_int * max(1, n))(*lengths)
    code = (C.c_uint16 * 256)()
    cl = (C.c_uint8 * 256)()
    rc = lib().orc_canonical_codes(s, ln, n, code, cl)
    if rc < 0:
        raise ValueError(f"canonical_codes rc={rc}")
    return np.array(code[:], dtype=np.uint16), np.array(cl[:], dtype=np.uint8)


class BitWriter:
    def __init__(self, flush_with_ones: bool, stuff: bool = False):
        self._bw = _BitWriter()
        lib().orc_bw_init(C.byref(self._bw), int(flush_with_ones), int(stuff))

    def write_bits(self, data: bytes, count: int):
        buf = (C.c_uint8 * max(1, len(data)))(*data)
        lib().orc_bw_write_bits(C.byref(self._bw), buf, count)

    def flush(self):
        lib().orc_bw_flush(C.byref(self._bw))

    def bytes(self) -> bytes:
        return bytes(self._bw.data[: self._bw.len])

    def __del__(self):
        try:
            lib().orc_bw_free(C.byref(self._bw))
        except Exception:
            pass


def stuff_bytes(data: bytes) -> bytes:
    src = (C.c_uint8 * max(1, len(data)))(*data)
    dst = (C.c_uint8 * (2 * len(data) + 1))()
    n = lib().orc_stuff_bytes(src, len(data), dst)
    return bytes(dst[:n])


PPM_ERRORS = {-1: "missing token", -2: "token parse failed", -3: "incomplete pixel",
              -4: "size mismatch", -5: "sample > max", -6: "not P3"}


def parse_ppm(text: bytes):
    """-> (w, h, max, samples u16 [h, w, 3]); raises ValueError(code) like the reference errors."""
    w, h, mx = C.c_int(), C.c_int(), C.c_int()
    p = C.POINTER(C.c_uint16)()
    rc = lib().orc_parse_ppm(text, len(text), C.byref(w), C.byref(h), C.byref(mx), C.byref(p))
    if rc < 0:
        raise ValueError(rc)
    n = w.value * h.value * 3
    arr = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint16)
    lib().orc_free(p)
    return w.value, h.value, mx.value, arr.reshape(h.value, w.value, 3)


# ------------------------------------------------------------------------------ whole path
@dataclass
class EncodeResult:
    jpeg: bytes
    width: int
    height: int
    padded_width: int
    padded_height: int
    y_blocks: int
    c_blocks: int
    n_mcus: int
    n_stream_blocks: int
    hist: np.ndarray            # [4, 256] u64: Y-DC, Y-AC, C-DC, C-AC
    tables: list                # 4 x (symbols, lengths) in ascending-frequency order
    scan_bits: int
    scan_bytes_unstuffed: int
    scan_bytes_stuffed: int
    header_bytes: int
    t_transform_s: float
    t_encode_s: float
    dct_y: np.ndarray | None = None     # [y_blocks, 64] f32 natural order, raster blocks
    dct_cb: np.ndarray | None = None
    dct_cr: np.ndarray | None = None
    stream: np.ndarray | None = None    # [n_stream_blocks, 64] i16 zig-zag, stream order
    extra: dict = field(default_factory=dict)


def _fmt_of(pixels: np.ndarray) -> int:
    if pixels.dtype == np.uint8:
        return FMT_U8
    if pixels.dtype == np.uint16:
        return FMT_U16
    if pixels.dtype == np.float32:
        return FMT_F32_NORM
    raise TypeError(pixels.dtype)


def encode(pixels: np.ndarray, max_value: int = 255, preset: int = P420, bits: int = 8,
           qpreset: int = 0, nthreads: int = 1, keep_planes: bool = False) -> EncodeResult:
    """pixels: [H, W, 3] u8 / u16 / f32(normalised)."""
    px = np.ascontiguousarray(pixels)
    h, w, c = px.shape
    assert c == 3
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    res = _Result()
    rc = lib().orc_encode(px.ctypes.data_as(C.c_void_p), _fmt_of(px), w, h, max_value, preset,
                          bits, qpreset, nthreads, C.byref(out), C.byref(n), C.byref(res),
                          int(keep_planes))
    if rc != 0:
        raise ValueError(f"orc_encode rc={rc}")
    jpeg = bytes(out[: n.value])
    lib().orc_free(out)
    hist = np.array([list(res.hist[t]) for t in range(4)], dtype=np.uint64)
    tables = [(list(res.table_sym[t][: res.table_n[t]]), list(res.table_len[t][: res.table_n[t]]))
              for t in range(4)]
    r = EncodeResult(jpeg, res.width, res.height, res.padded_width, res.padded_height,
                     res.y_blocks, res.c_blocks, res.n_mcus, res.n_stream_blocks, hist, tables,
                     res.scan_bits, res.scan_bytes_unstuffed, res.scan_bytes_stuffed,
                     res.header_bytes, res.t_transform_s, res.t_encode_s)
    if keep_planes:
        r.dct_y = np.ctypeslib.as_array(res.dct_y, shape=(res.y_blocks, 64)).copy()
        r.dct_cb = np.ctypeslib.as_array(res.dct_cb, shape=(res.c_blocks, 64)).copy()
        r.dct_cr = np.ctypeslib.as_array(res.dct_cr, shape=(res.c_blocks, 64)).copy()
        r.stream = np.ctypeslib.as_array(res.stream, shape=(res.n_stream_blocks, 64)).copy()
        lib().orc_result_free(C.byref(res))
    return r


def encode_ppm(text: bytes, preset: int = P420, bits: int = 8, qpreset: int = 0,
               nthreads: int = 1, keep_planes: bool = False) -> EncodeResult:
    w, h, mx, samples = parse_ppm(text)
    px = samples.astype(np.uint8) if mx <= 255 else samples
    return encode(px, mx, preset, bits, qpreset, nthreads, keep_planes)
