"""Writes tests/golden/config5_sha256.json: SHA-256 and length of the ORACLE's output for BASELINE config 5 at full
size -- one 32768 x 32768 image, synth.smooth(5) rows generated in 2048-row slabs exactly as
tests/test_cuda_parity.py::test_config5_full_size_shards_equal_single_chain generates them on the device (the
generator is integer-hash based: CPU and CUDA tensors are bit-identical, tests/test_host_logic.py checks a slab).
The oracle needs ~30 GB of host memory and a few minutes for this size, so the GPU test compares against this
committed digest instead of running it.

  python tests/golden/make_config5_sha.py [--size 32768]
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dmmt_jpeg_encoder_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    n = a.size
    px = np.empty((n, n, 3), np.uint8)
    for y0 in range(0, n, 2048):
        px[y0:y0 + 2048] = synth.make("smooth", 5, min(2048, n - y0), n, "cpu", y0=y0).numpy()
    t = time.time()
    r = O.encode(px, 255, O.P420, 8, 0, nthreads=a.threads)
    out = {"width": n, "height": n, "kind": "smooth", "index": 5, "slab_rows": 2048, "preset": "P420", "qpreset": 0,
           "sha256": hashlib.sha256(r.jpeg).hexdigest(), "bytes": len(r.jpeg), "scan_bits": int(r.scan_bits),
           "oracle_seconds": round(time.time() - t, 1)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config5_sha256.json")
    known = {}
    if os.path.exists(path):
        known = json.load(open(path))
    known[str(n)] = out
    json.dump(known, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
