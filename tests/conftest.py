import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def load_fixture(name: str):
    """-> (ppm_text_bytes or None, pixels u8 [H,W,3], max).  500x500 is stored as PNG and its
    P3 text is re-emitted GIMP-style (comment on line 2, one token per line) for the parser."""
    if name == "500x500":
        from PIL import Image

        px = np.array(Image.open(os.path.join(GOLDEN, "inputs", "500x500.png")).convert("RGB"))
        lines = ["P3", "# Created by GIMP version 2.10.34 PNM plug-in", "500 500", "255"]
        lines += [str(v) for v in px.reshape(-1)]
        return ("\n".join(lines) + "\n").encode(), px, 255
    text = open(os.path.join(GOLDEN, "inputs", name + ".ppm"), "rb").read()
    from oracle import oracle as O

    w, h, mx, s = O.parse_ppm(text)
    return text, s.astype(np.uint8), mx


FIXTURES = ["small", "8x8", "16x16", "7x17", "500x500"]
PRESETS = {"P444": 0, "P422": 1, "P420": 2}


def synth_image(kind: str, w: int, h: int, seed: int = 0) -> np.ndarray:
    """Deterministic synthetic generators of SURVEY.md section 8d (grad / photo / uniform)."""
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    if kind == "grad":
        r = (x + 8 * y) & 255
        g = (2 * x + 3 * y + 85) & 255
        b = (5 * x + y + 170) & 255
        return np.stack([r, g, b], -1).astype(np.uint8)
    rng = np.random.default_rng(1234 + seed)
    if kind == "uniform":
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    if kind == "photo":
        n = rng.standard_normal((h, w, 3)).astype(np.float32)
        for _ in range(3):
            n = (n + np.roll(n, 1, 0) + np.roll(n, -1, 0) + np.roll(n, 1, 1) + np.roll(n, -1, 1)) / 5
        n = n / (n.std() + 1e-6)
        ramp = ((x + y) / float(w + h))[..., None].astype(np.float32)
        img = 0.6 * (0.5 + 0.25 * n) + 0.4 * ramp + rng.normal(0, 2 / 255, (h, w, 3))
        return np.clip(np.rint(img * 255), 0, 255).astype(np.uint8)
    raise ValueError(kind)
