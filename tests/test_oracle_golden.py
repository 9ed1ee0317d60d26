"""Whole-file checks of the oracle: committed golden JPEGs + the survey's independent SHA-256
table (tests/golden/expected_sha256.json), libjpeg (PIL) decodability, stage consistency."""
import hashlib
import io
import json
import os

import numpy as np
import pytest

from conftest import FIXTURES, GOLDEN, PRESETS, load_fixture, synth_image
from oracle import oracle as O

SHA = json.load(open(os.path.join(GOLDEN, "expected_sha256.json")))


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("pname", list(PRESETS))
def test_oracle_reproduces_golden(name, pname):
    text, px, mx = load_fixture(name)
    r = O.encode_ppm(text, preset=PRESETS[pname])
    key = f"{name}_{pname}"
    assert hashlib.sha256(r.jpeg).hexdigest() == SHA[key]
    assert r.jpeg == open(os.path.join(GOLDEN, "jpeg", key + ".jpg"), "rb").read()
    # the u8 pixel path gives the same file as the parsed P3 text
    assert O.encode(px, mx, PRESETS[pname]).jpeg == r.jpeg


@pytest.mark.parametrize("pname", list(PRESETS))
def test_decodes_with_libjpeg(pname):
    from PIL import Image

    y, x = np.mgrid[0:80, 0:96]
    px = np.stack([x * 2, y * 3, x + y], -1).astype(np.uint8)   # smooth: survives 4:2:0
    r = O.encode(px, 255, PRESETS[pname])
    im = np.array(Image.open(io.BytesIO(r.jpeg)).convert("RGB")).astype(np.float64)
    mse = ((im - px.astype(np.float64)) ** 2).mean()
    assert 10 * np.log10(255 ** 2 / mse) > 38.0


def test_input_formats_agree():
    px = synth_image("uniform", 40, 24, 3)
    a = O.encode(px, 255, O.P420).jpeg
    b = O.encode(px.astype(np.uint16), 255, O.P420).jpeg
    c = O.encode((px.astype(np.float32) / np.float32(255)), 255, O.P420).jpeg
    assert a == b == c


def test_threads_do_not_change_output():
    px = synth_image("photo", 256, 128, 2)
    assert O.encode(px, 255, O.P420, nthreads=1).jpeg == O.encode(px, 255, O.P420, nthreads=4).jpeg


def test_stage_consistency_and_length16():
    """hist x code lengths == scan bits; uniform noise drives AC code lengths to 16 (15 + the
    '+1' quirk of symbol_counting.rs:88), which must be legal."""
    px = synth_image("uniform", 256, 256, 0)
    r = O.encode(px, 255, O.P420, keep_planes=True)
    bits = 0
    for t, (sym, ln) in enumerate(r.tables):
        for s, l in zip(sym, ln):
            bits += int(r.hist[t][s]) * (l + (s & 15))
    assert bits == r.scan_bits
    assert max(max(ln) for _, ln in r.tables) <= 16
    assert r.scan_bytes_unstuffed == (r.scan_bits + 7) // 8
    assert len(r.jpeg) == r.header_bytes + r.scan_bytes_stuffed + 2
    # DC of the quantised stream equals round(dct/q) of the block-contiguous planes
    q = O.qtable(0, False)
    assert r.stream[0, 0] == O.quantize(float(r.dct_y[0, 0]), int(q[0]))


@pytest.mark.parametrize("q", range(7))
def test_all_quant_presets_run(q):
    px = synth_image("grad", 33, 17)
    j = O.encode(px, 255, O.P420, qpreset=q).jpeg
    assert j[25:25 + 64] == bytes(O.qtable(q, False)[O.zigzag()])


def test_config5_digest_script_is_reproducible_at_4096():
    """tests/golden/config5_sha256.json (oracle digests of the config-5 workload, made by make_config5_sha.py): the
    4096 x 4096 entry is recomputed here, which pins generator + oracle + script for the 32768 x 32768 entry that the
    GPU test compares the CUDA file with."""
    from dmmt_jpeg_encoder_b200 import synth

    gold = json.load(open(os.path.join(GOLDEN, "config5_sha256.json")))
    g = gold["4096"]
    n = g["width"]
    px = np.empty((n, n, 3), np.uint8)
    for y0 in range(0, n, g["slab_rows"]):
        px[y0:y0 + g["slab_rows"]] = synth.make(g["kind"], g["index"], g["slab_rows"], n, "cpu", y0=y0).numpy()
    r = O.encode(px, 255, O.P420, 8, 0, nthreads=os.cpu_count() or 1)
    assert len(r.jpeg) == g["bytes"] and hashlib.sha256(r.jpeg).hexdigest() == g["sha256"]
    assert set(gold["32768"]) == set(g) and gold["32768"]["bytes"] > 100e6
