"""Pins the CPU oracle against the reference's own unit-test known-answer vectors
(SURVEY.md section 8c).  Each test cites the reference test (file:line under /root/reference/src)."""
import numpy as np
import pytest

from oracle import oracle as O


# ---- color.rs:106-209
def test_rgb_to_ycbcr_kat():
    y, cb, cr = O.rgb_to_ycbcr(0.25, 0.75, 0.333)          # color.rs:106-129
    assert 12.95 <= y < 13.05 and -31.68 <= cb < -31.58 and -55.13 <= cr < -55.03
    y, cb, cr = O.rgb_to_ycbcr(1.0, 1.0, 1.0)              # color.rs:131-154
    assert 126.99999 <= y <= 127.00001 and abs(cb) <= 0.5 and abs(cr) <= 0.5
    y, cb, cr = O.rgb_to_ycbcr(0.0, 0.0, 0.0)              # color.rs:156-167 (exact)
    assert (y, cb, cr) == (-128.0, 0.0, 0.0)


def test_normalize_kat():
    assert 7.209e-3 <= O.normalize(128, 17734) <= 7.219e-3  # color.rs:169-185
    assert O.normalize(14355, 17734) >= 0.809459
    assert 4.99e-4 <= O.normalize(9, 17734) <= 5.09e-4
    assert O.normalize(65535, 65535) == 1.0                 # color.rs:187-194
    assert 0.133333 <= O.normalize(2, 15) <= 0.133334       # color.rs:196-209
    assert 0.333333 <= O.normalize(5, 15) <= 0.333334
    assert O.normalize(15, 15) == 1.0


# ---- padder.rs:52-87
def test_padding_kat():
    assert O.padded_dims(1, 1, O.P422) == (16, 8)           # pad_one: 16 x 8
    assert O.padded_dims(17, 7, O.P420) == (32, 16)         # pad_7_17 @ (16,16)
    assert O.padded_dims(7, 17, O.P420) == (16, 32)
    assert O.padded_dims(500, 500, O.P444) == (504, 504)


# ---- arai.rs:117-219
TEST_VALUES = np.array([
    1, 2, 1, 2, 3, 2, 3, 2, 3, 2, 1, 2, 3, 4, 3, 2, 3, 4, 3, 2, 3, 4, 5, 6, 7, 6, 5, 4, 3, 2, 3, 2,
    3, 4, 5, 5, 6, 5, 2, 3, 4, 3, 2, 3, 4, 5, 4, 3, 2, 3, 4, 5, 6, 5, 4, 3, 2, 3, 4, 5, 3, 4, 3, 4,
], dtype=np.float32)
F = np.float32
A1 = F(0.70710678118654752440); A2 = F(0.5411961); A3 = A1; A4 = F(1.3065629); A5 = F(0.3826834)
S = [F(0.3535533), F(0.2548978), F(0.27059805), F(0.30067244), F(0.35355338), F(0.4499881),
     F(0.6532815), F(1.2814577)]


def test_arai_1d_bit_equal_closed_forms():
    """arai.rs:117-165,204-219 compare_fast_own: the 1-D pass is BIT-equal to the closed forms
    (evaluated here in numpy float32, left to right like the Rust expressions)."""
    i = TEST_VALUES[:8]
    out = O.fast_arai(i)
    s = F(0)
    for v in i:
        s = F(s + v)
    assert out[0] == F(s * S[0])
    y4 = F(F(F(F(F(F(F(i[0] + i[7]) + i[3]) + i[4]) - i[1]) - i[6]) - i[2]) - i[5])
    assert out[4] == F(y4 * S[4])
    t = F(F(F(F(F(F(F(i[0] + i[1]) - i[2]) - i[3]) - i[4]) - i[5]) + i[6]) + i[7])
    y2 = F(F(F(F(F(t * A1) + i[0]) + i[7]) - i[3]) - i[4])
    assert out[2] == F(y2 * S[2])
    y6 = F(F(F(F(F(t * -A1) + i[0]) + i[7]) - i[3]) - i[4])
    assert out[6] == F(y6 * S[6])
    a = F(F(F(i[3] - i[4]) + i[2]) - i[5])
    b = F(F(F(F(a - i[1]) + i[6]) - i[0]) + i[7])
    c = F(F(F(i[2] - i[5]) + i[1]) - i[6])
    y5 = F(F(F(F(F(A2 * a) + F(A5 * b)) + i[0]) - i[7]) - F(A3 * c))
    assert out[5] == F(y5 * S[5])
    temp = F(F(F(i[1] - i[6]) + i[0]) - i[7])
    d = F(F(F(F(temp - i[3]) + i[4]) - i[0]) + i[7])
    y1 = F(F(F(F(i[0] - i[7]) + F(A3 * c)) + F(A4 * temp)) - F(A5 * d))
    assert out[1] == F(y1 * S[1])
    y7 = F(F(F(F(i[0] - i[7]) + F(A3 * c)) - F(A4 * temp)) + F(A5 * d))
    assert out[7] == F(y7 * S[7])
    y3 = F(F(F(F(F(-A2 * a) - F(A5 * b)) + i[0]) - i[7]) - F(A3 * c))
    assert out[3] == F(y3 * S[3])


def _direct_dct(block):
    """simple.rs:19-99 direct O(N^4) orthonormal DCT-II (float64)."""
    x = block.reshape(8, 8).astype(np.float64)
    n = np.arange(8)
    Cm = np.cos((2 * n[None, :] + 1) * n[:, None] * np.pi / 16)
    c = np.where(n == 0, 1 / np.sqrt(2), 1.0)
    return (0.25 * (c[:, None] * c[None, :]) * (Cm @ x @ Cm.T)).reshape(64)


def test_arai_vs_direct_dct():
    """arai.rs:190-201 test_fast_simple: <= 1e-4 absolute on TEST_VALUES."""
    got = O.dct8x8(TEST_VALUES)
    assert np.abs(got - _direct_dct(TEST_VALUES)).max() <= 1e-4
    rng = np.random.default_rng(7)
    for _ in range(20):
        b = rng.uniform(-128, 127, 64).astype(np.float32)
        ref = _direct_dct(b)
        assert np.abs(O.dct8x8(b) - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


# ---- frequency_block.rs:67-100
def test_zigzag_kat():
    block = np.array([0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30,
                      41, 43, 9, 11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22,
                      33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58,
                      62, 63])
    assert list(block[O.zigzag()]) == list(range(64))


# ---- block_entangler.rs:99-165
def test_quadfold_kat():
    seq = [0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15]
    out = [seq[8 * (i // 8) + O.quadfold_index(i % 8, 4)] for i in range(16)]
    assert out == list(range(16))
    seq = [0, 1, 4, 5, 8, 9, 2, 3, 6, 7, 10, 11, 12, 13, 16, 17, 20, 21, 14, 15, 18, 19, 22, 23]
    out = [seq[12 * (i // 12) + O.quadfold_index(i % 12, 6)] for i in range(24)]
    assert out == list(range(24))


# ---- categorize.rs:175-289
def test_categorize_kat():
    assert O.categorize(57) == (6, 0b1110010000000000)
    assert O.categorize(45) == (6, 0b1011010000000000)
    assert O.categorize(1) == (1, 0b1000000000000000)
    assert O.categorize(-30) == (5, 0b0000100000000000)
    assert O.categorize(32767) == (15, 0b1111111111111110)
    assert O.categorize(-32767) == (15, 0)
    assert O.categorize(0) == (0, 0)
    assert O.categorize(-32768)[0] == -1   # reference panics (categorize.rs:236-240)


def test_rle_tokens_kat():
    seq = [57, 45, 0, 0, 0, 0, 23, 0, -30, -16] + [0] * 19 + [1, 0]
    assert O.rle_tokens(seq) == [(0, 57), (0, 45), (4, 23), (1, -30), (0, -16), (15, 0), (3, 1),
                                 (0, 0)]
    assert O.rle_tokens([0] * 63) == [(0, 0)]
    assert O.rle_tokens([0] * 62 + [5]) == [(15, 0), (15, 0), (15, 0), (14, 5)]


# ---- symbol_counting.rs:108-198 test_count_symbols, through the oracle's own counting code
# (orc_count_block, the function orc_encode's histogram loop calls) and its frequency sort
def test_count_symbols_kat():
    blocks = [(30, [(0, 300), (15, 0), (4, 5), (0, 0)]), (0, [(0, 600), (15, 0), (4, 15), (0, 0)]),
              (60, [(0, 100), (15, 0), (2, 7), (0, 0)]), (1, [(0, 900), (15, 0), (0, 1), (0, 0)])]
    dc, ac = O.count_blocks(blocks)
    assert {i: int(c) for i, c in enumerate(dc) if c} == {5: 1, 0: 1, 6: 1, 1: 1}
    assert {i: int(c) for i, c in enumerate(ac) if c} == {
        0b00001001: 1, 0b11110000: 4, 0b01000011: 1, 0b00000000: 4, 0b00001010: 2, 0b01000100: 1,
        0b00000111: 1, 0b00100011: 1, 0b00000001: 1}
    # to_symbol_frequencies + sort_by_frequency (:25-32, :92-94): ascending symbol, stable by frequency
    sym, _ = O.build_table(ac)
    assert sym == [0b00000001, 0b00000111, 0b00001001, 0b00100011, 0b01000011, 0b01000100, 0b00001010,
                   0b00000000, 0b11110000]
    with pytest.raises(ValueError):
        O.count_blocks([(-32768, [])])   # categorize.rs:236-240 panics


# ---- subsampling.rs:332-550
CH1 = np.arange(1, 17, dtype=np.float32).reshape(4, 4)     # TEST_CHANNEL_ONE
CH2 = np.arange(1, 65, dtype=np.float32).reshape(8, 8)     # TEST_CHANNEL_TWO


def test_subsampling_value_kats():
    assert O.subsample_value(CH1, 1, 1, False, 2, 1) == 7.0      # no_subsampling_test :342-362
    assert O.subsample_value(CH1, 2, 1, False, 1, 1) == 7.0      # skip_subsampling_test :364-384
    assert O.subsample_value(CH1, 1, 2, True, 1, 1) == 12.0      # average_subsampling_test :386-406
    assert O.subsample_value(CH1, 2, 1, True, 2, 2) is None      # out_of_bounds_high :408-424
    assert O.subsample_value(CH1, 2, 3, True, 1, 1) == 15.5      # repeat_border_test (clamp) :426-446
    # the summation order of the 2x2 window is (x,y),(x,y+1),(x+1,y),(x+1,y+1) (rect(): x outer, :116-121):
    # values chosen so that every other association of the four f32 additions rounds differently
    w = np.array([[1.0, 2.0 ** -24], [2.0 ** -24, 2.0 ** -24]], np.float32)   # (y, x) layout
    got = O.subsample_value(w, 2, 2, True, 0, 0)
    s = np.float32(0.0)
    for x in range(2):
        for y in range(2):
            s = np.float32(s + w[y, x])
    assert got == np.float32(s / np.float32(4.0))
    other = np.float32(np.float32(np.float32(w[1, 1] + w[1, 0]) + w[0, 1]) + w[0, 0]) / np.float32(4.0)
    assert got != other


def test_square_resorter_kats():
    assert O.subsample_retile(CH1, 1, 1, False, 4).tolist() == CH1.reshape(-1).tolist()   # :448-466
    exp_1x1 = [1, 2, 3, 4, 9, 10, 11, 12, 17, 18, 19, 20, 25, 26, 27, 28, 5, 6, 7, 8, 13, 14, 15, 16, 21, 22,
               23, 24, 29, 30, 31, 32, 33, 34, 35, 36, 41, 42, 43, 44, 49, 50, 51, 52, 57, 58, 59, 60, 37, 38,
               39, 40, 45, 46, 47, 48, 53, 54, 55, 56, 61, 62, 63, 64]
    assert O.subsample_retile(CH2, 1, 1, False, 4).tolist() == exp_1x1                       # :468-497
    exp_2x2 = [1, 3, 5, 7, 17, 19, 21, 23, 33, 35, 37, 39, 49, 51, 53, 55]
    assert O.subsample_retile(CH2, 2, 2, False, 4).tolist() == exp_2x2                       # :499-523
    exp_1x2 = [1, 2, 3, 4, 17, 18, 19, 20, 33, 34, 35, 36, 49, 50, 51, 52, 5, 6, 7, 8, 21, 22, 23, 24, 37, 38,
               39, 40, 53, 54, 55, 56]
    assert O.subsample_retile(CH2, 1, 2, False, 4).tolist() == exp_1x2                       # :525-550


def test_encode_uses_the_pinned_retile_and_count():
    """The whole-path oracle goes through the same functions the KATs above pin: its chroma planes equal
    orc_subsample_retile of the colour-converted padded image, its histogram equals orc_count_block's."""
    rng = np.random.default_rng(5)
    px = rng.integers(0, 256, (19, 27, 3), dtype=np.uint8)
    r = O.encode(px, 255, O.P420, keep_planes=True)
    n = px.astype(np.float32) / np.float32(255.0)
    pad = O.pad_image(n, 16, 16)
    ycc = np.array([[O.rgb_to_ycbcr(*p) for p in row] for row in pad], np.float32)
    cb = O.subsample_retile(ycc[..., 1], 2, 2, True, 8).reshape(-1, 64)
    tmp = cb.copy()
    for b in tmp:
        b[:] = O.dct8x8(b)
    assert np.array_equal(tmp.view(np.uint32) & 0x7FFFFFFF, r.dct_cb.view(np.uint32) & 0x7FFFFFFF)
    blocks, last = [], [0, 0, 0]
    for i, blk in enumerate(r.stream):
        comp = 0 if i % 6 < 4 else i % 6 - 3
        d = int(np.int16(int(blk[0]) - last[comp]))
        last[comp] = int(blk[0])
        blocks.append((comp, d, O.rle_tokens(blk[1:])))
    for sel, t in ((lambda c: c == 0, 0), (lambda c: c > 0, 2)):
        dc, ac = O.count_blocks([(d, tk) for c, d, tk in blocks if sel(c)])
        assert np.array_equal(dc, r.hist[t][:16]) and np.array_equal(ac, r.hist[t + 1])


# ---- padder.rs:52-87 (sizes) and :18-38 (content: source dots, black to the right and below)
def test_padder_kats():
    red = np.array([1.0, 0.0, 0.0], np.float32)
    p = O.pad_image(np.tile(red, (1, 1, 1)), 16, 8)                      # pad_one
    assert p.shape == (8, 16, 3) and p.shape[0] * p.shape[1] == 16 * 8
    assert p[0, 0].tolist() == [1.0, 0.0, 0.0] and not p[0, 1:].any() and not p[1:].any()
    p = O.pad_image(np.tile(red, (7, 17, 1)), 16, 16)                    # pad_7_17
    assert p.shape == (16, 32, 3) and p.shape[0] * p.shape[1] == 32 * 16
    assert (p[:7, :17] == red).all() and not p[:7, 17:].any() and not p[7:].any()
    p = O.pad_image(np.tile(red, (99, 99, 1)), 10, 10)                   # pad_99_99
    assert p.shape[0] * p.shape[1] == 10000
    assert (p[:99, :99] == red).all() and not p[:, 99].any() and not p[99].any()
    # black pads to YCbCr (-128, 0, 0) exactly (color.rs:157-167), which is what the DCT sees there
    assert O.rgb_to_ycbcr(*p[99, 99]).tolist() == [-128.0, 0.0, 0.0]


# ---- length_limited.rs:209-264, tree.rs:349-405
def test_package_merge_kats():
    assert O.package_merge([1, 2, 5, 8, 10, 11, 14, 14, 15, 18, 20], 4) == [4] * 6 + [3] * 5
    assert O.package_merge([1, 1, 1, 2, 2, 2, 3, 6, 17, 20], 5) == [5, 5, 4, 4, 4, 4, 4, 3, 2, 2]
    assert O.package_merge([1, 1, 1, 2, 2, 2, 3, 6, 17, 20], 4) == [4] * 8 + [2, 2]
    with pytest.raises(ValueError):
        O.package_merge([1, 1, 1, 2, 2, 2, 3, 6, 17, 20], 3)
    # tree.rs:344-372: depths for limit 10 (sorted by frequency)
    # test helper depth = code length + 1 (root counts as depth 1): [5,5,4,3,3,3] and [5,5,4,4,4,3,3]
    assert O.package_merge(sorted([17, 3, 12, 3, 18, 12]), 10) == [4, 4, 3, 2, 2, 2]
    assert O.package_merge(sorted([17, 3, 12, 3, 18, 12, 13]), 10) == [4, 4, 3, 3, 3, 2, 2]
    hist = np.zeros(256, np.uint64)
    for s, f in [(1, 17), (2, 3), (3, 12), (4, 3), (5, 18), (6, 12), (7, 13)]:
        hist[s] = f
    sym, ln = O.build_table(hist, limit=10, plus_one=True)
    # tree.rs:391-405 one-star replaced depths [6,5,4,4,4,3,3] == lengths with the +1 quirk + 1
    # (tree depth counts the root); the code-length vector itself is:
    assert sym == [2, 4, 3, 6, 7, 1, 5]
    assert [x + 1 for x in ln] == [6, 5, 4, 4, 4, 3, 3]
    assert O.package_merge([5], 15) == [0]            # n = 1 -> length 0, +1 quirk makes it 1


# ---- huffman/encoder.rs:213-302
SYMS_FREQS = [(1, 14), (2, 30), (3, 4), (4, 7), (5, 9), (6, 4), (7, 42), (8, 1), (9, 14), (10, 5),
              (11, 14), (12, 30), (13, 4), (14, 7), (15, 9), (16, 4), (17, 42), (18, 1), (19, 14),
              (20, 5), (21, 14), (22, 30), (23, 4), (24, 7), (25, 9), (26, 4), (27, 42), (28, 1),
              (29, 14), (30, 12), (31, 32), (32, 1)]
SYMBOL_SEQ = [27, 17, 7, 31, 22, 12, 2, 29, 21, 19, 11, 9, 1, 30, 25, 15, 5, 24, 14, 4, 20, 10, 26,
              23, 16, 13, 6, 3, 32, 28, 18, 8]
BYTE_SEQ = bytes([0b00000100, 0b01101000, 0b10101100, 0b11110000, 0b10001100, 0b10100111,
                  0b01001010, 0b11011010, 0b11101011, 0b11110000, 0b11000111, 0b00101100,
                  0b11110100, 0b11010111, 0b01101101, 0b11111000, 0b11100111, 0b10101110,
                  0b11111100, 0b11110111, 0b11101111, 0b11000000])


def test_coder_encode_22_bytes():
    """huffman/encoder.rs:241-269 test_coder_encode: tables + canonical codes + bit packing
    jointly -> 22 exact bytes (limit 6, +1 quirk, zero padding)."""
    hist = np.zeros(256, np.uint64)
    for s, f in SYMS_FREQS:
        hist[s] = f
    sym, ln = O.build_table(hist, limit=6, plus_one=True)
    code, clen = O.canonical_codes(sym, ln)
    bw = O.BitWriter(flush_with_ones=False)
    for s in SYMBOL_SEQ:
        c = int(code[s])
        bw.write_bits(bytes([c >> 8, c & 255]), int(clen[s]))
    bw.flush()
    assert bw.bytes() == BYTE_SEQ


def test_canonical_validation():
    with pytest.raises(ValueError):
        O.canonical_codes([0, 1, 2, 3], [1, 5, 4, 3])     # unsorted  (encoder.rs:199-204)
    with pytest.raises(ValueError):
        O.canonical_codes([0, 1, 2, 3], [17, 5, 4, 3])    # > 16 bits (encoder.rs:206-211)
    code, clen = O.canonical_codes([3, 1], [5, 1])
    assert (code[1], clen[1], code[3], clen[3]) == (0, 1, 0x8000, 5)


# ---- binary_stream.rs:104-158
def test_bitwriter_kats():
    bw = O.BitWriter(False)
    bw.write_bits(bytes([72, 65, 76, 76, 79]), 40)
    bw.flush()
    assert bw.bytes() == bytes([72, 65, 76, 76, 79])
    bw = O.BitWriter(False)
    for b, n in [(0xFF, 2), (0x00, 4), (0xFF, 2), (0xFF, 4)]:
        bw.write_bits(bytes([b]), n)
    bw.flush()
    assert bw.bytes() == bytes([195, 15 << 4])
    bw = O.BitWriter(False)
    bw.write_bits(bytes([0xFF]), 3)
    bw.write_bits(bytes([1, 2, 4 | 128]), 24)
    bw.flush()
    assert bw.bytes() == bytes([224, 32, 80, 128])
    bw = O.BitWriter(True)
    bw.write_bits(bytes([0]), 3)
    bw.flush()
    assert bw.bytes() == bytes([31])


# ---- segment_marker_injector.rs:43-58
def test_stuffing_kat():
    assert O.stuff_bytes(bytes([1, 2, 0xFF, 0, 3])) == bytes([1, 2, 0xFF, 0, 0, 3])


# ---- jpeg/encoder.rs:449-577 (through a whole file: segments at their fixed offsets)
def test_header_segments_kat():
    px = np.zeros((2, 3, 3), np.uint8)
    j = O.encode(px, 255, O.P444).jpeg
    assert j[:2] == b"\xff\xd8"
    assert j[2:20] == bytes([0xFF, 0xE0, 0, 0x10]) + b"JFIF\0" + bytes([1, 2, 0, 0, 0x48, 0, 0x48, 0, 0])
    dqt = bytes([0xFF, 0xDB, 0x00, 0x43, 0x00, 16, 11, 12, 14, 12, 10, 16, 14, 13, 14, 18, 17, 16, 19,
                 24, 40, 26, 24, 22, 22, 24, 49, 35, 37, 29, 40, 58, 51, 61, 60, 57, 51, 56, 55, 64,
                 72, 92, 78, 64, 68, 87, 69, 55, 56, 80, 109, 81, 87, 95, 98, 103, 104, 103, 62, 77,
                 113, 121, 112, 100, 120, 92, 101, 103, 99])
    assert j[20:89] == dqt
    assert j[89:94] == bytes([0xFF, 0xDB, 0x00, 0x43, 0x01])
    sof = bytes([0xFF, 0xC0, 0x00, 0x11, 8, 0, 2, 0, 3, 3, 1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1])
    assert j[158:177] == sof
    sos = bytes([0xFF, 0xDA, 0x00, 0x0C, 3, 1, 0x01, 2, 0x23, 3, 0x23, 0, 0x3F, 0])
    assert sos in j and j[-2:] == b"\xff\xd9"
    assert O.encode(px, 255, O.P422).jpeg[158 + 11] == 0x21
    assert O.encode(px, 255, O.P420).jpeg[158 + 11] == 0x22


# ---- image/reader/ppm.rs:253-307
def test_ppm_parser():
    w, h, mx, s = O.parse_ppm(b"P3\n2 1 # c\n255\n1 2 3 4 5 6")
    assert (w, h, mx) == (2, 1, 255) and s.reshape(-1).tolist() == [1, 2, 3, 4, 5, 6]
    for text, code in [(b"P6 1 1 255 1 2 3", -6), (b"P3 1 1 255 1 2", -3), (b"P3 2 1 255 1 2 3", -4),
                       (b"P3 1 1 255 1 2 x", -2), (b"P3 1 1", -1), (b"P3 1 1 9 1 2 10", -5),
                       (b"", -1), (b"P3 70000 1 255", -2)]:
        with pytest.raises(ValueError) as e:
            O.parse_ppm(text)
        assert e.value.args[0] == code
    # '#' inside a token does not split it (ppm.rs:49-67)
    assert O.parse_ppm(b"P3 1 1 255 1#x\n2 3 4")[3].reshape(-1).tolist() == [12, 3, 4]


def test_quantize_round_half_away():
    assert O.quantize(8.0, 16) == 1 and O.quantize(-8.0, 16) == -1      # ties away from zero
    assert O.quantize(7.9999995, 16) == 0 and O.quantize(24.0, 16) == 2
    assert O.quantize(1e9, 1) == 32767 and O.quantize(-1e9, 1) == -32768
