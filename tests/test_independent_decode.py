"""Independent check of the bitstream: a T.81 entropy decoder (tests/jpeg_entropy_decoder.py, written from
the standard, not from the reference) must read back EXACTLY the oracle's quantised coefficients, tables
and geometry from the files -- for the committed golden files (which the CUDA path reproduces byte for
byte) and for fresh oracle output.  This pins what the reference's KATs leave open: that the DHT/DQT/SOF/
SOS bytes, the canonical code assignment (huffman/encoder.rs:37-157), the 16-bit codes of the '+1' quirk,
ZRL/EOB handling, the DC chains and the 1-padding mean the same thing to a standard decoder."""
import os

import numpy as np
import pytest

from conftest import FIXTURES, GOLDEN, PRESETS, load_fixture, synth_image
from jpeg_entropy_decoder import decode
from oracle import oracle as O

SAMPLING = {0: [(1, 1)] * 3, 1: [(2, 1), (1, 1), (1, 1)], 2: [(2, 2), (1, 1), (1, 1)]}


def check(jpeg: bytes, r, preset: int, q: int, bits: int = 8):
    d = decode(jpeg)
    assert (d["width"], d["height"], d["precision"]) == (r.width, r.height, bits)
    assert d["sampling"] == SAMPLING[preset] and d["tq"] == [0, 1, 1]
    zz = O.zigzag()
    assert d["qtables"][0] == list(O.qtable(q, False)[zz]) and d["qtables"][1] == list(O.qtable(q, True)[zz])
    # marker order of encoder.rs:125-135: APP0, DQT, DQT, SOF0, DHT x4, SOS
    assert d["markers"] == [0xE0, 0xDB, 0xDB, 0xC0, 0xC4, 0xC4, 0xC4, 0xC4, 0xDA]
    assert sorted(d["dht"]) == [(0, 0), (0, 2), (1, 1), (1, 3)]          # ids of encoder.rs:78-84
    np.testing.assert_array_equal(d["coefficients"], r.stream.astype(np.int32))


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("pname", list(PRESETS))
def test_golden_files_decode_to_the_oracle_coefficients(name, pname):
    _, px, mx = load_fixture(name)
    r = O.encode(px, mx, PRESETS[pname], keep_planes=True)
    golden = open(os.path.join(GOLDEN, "jpeg", f"{name}_{pname}.jpg"), "rb").read()
    assert golden == r.jpeg
    check(golden, r, PRESETS[pname], 0)


@pytest.mark.parametrize("kind,w,h,preset,q", [("uniform", 96, 80, 0, 1), ("photo", 130, 75, 2, 0), ("grad", 77, 33, 1, 4),
                                               ("uniform", 64, 64, 2, 6), ("photo", 256, 256, 0, 0)])
def test_synthetic_files_decode_to_the_oracle_coefficients(kind, w, h, preset, q):
    """includes noise (16-bit codes, many ZRLs, stuffed 0xFF bytes) and every subsampling preset"""
    px = synth_image(kind, w, h, 11)
    r = O.encode(px, 255, preset, 8, q, keep_planes=True)
    check(r.jpeg, r, preset, q)


def test_sparse_blocks_exercise_zrl_and_eob():
    y, x = np.mgrid[0:64, 0:64]
    # the (7,7) DCT basis function: the only non-zero AC coefficient is number 63 -> three ZRLs, no EOB
    v = 128 + 60 * np.cos((2 * (x % 8) + 1) * 7 * np.pi / 16) * np.cos((2 * (y % 8) + 1) * 7 * np.pi / 16)
    px = np.clip(np.rint(v), 0, 255).astype(np.uint8)[..., None].repeat(3, -1)
    px[16:32, 16:48] = 40                                                   # plus flat blocks (EOB right after DC)
    r = O.encode(px, 255, O.P444, 8, 1, keep_planes=True)
    assert r.stream[0][63] != 0 and not r.stream[0][1:63].any()
    assert r.hist[1][0xF0] > 0                   # ZRL symbols really occur
    check(r.jpeg, r, O.P444, 1)
