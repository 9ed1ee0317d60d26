"""BASELINE config 5 through the C ABI alone: dmmt_encode_sharded (ONE process, one context per GPU, exchanges over
peer memory, K4 of every shard into the file on the first device).  The call takes host pixels and returns a host file;
what is reported is the phase section alone (DMMT_SHARDED_TIMING=1 -> dmmt_encode_sharded_last_ms: wall clock with
every H2D copy finished before it and its one host synchronisation at the end), comparable with bench.py's
extra.config5, plus the whole call.  The file is checked against the oracle's committed SHA-256.
  python tools/bench_c_abi_sharded.py [--size 32768] [--gpus N]
"""
import argparse
import hashlib
import json
import os
import sys
import time

os.environ["DMMT_SHARDED_TIMING"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmmt_jpeg_encoder_b200 as D  # noqa: E402
from dmmt_jpeg_encoder_b200 import _ffi as F  # noqa: E402
from dmmt_jpeg_encoder_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    n = a.size
    px = np.empty((n, n, 3), np.uint8)                       # the image in host memory, generated on device 0 in slabs
    for y0 in range(0, n, 2048):
        y1 = min(n, y0 + 2048)
        px[y0:y1] = synth.make("smooth", 5, y1 - y0, n, "cuda:0", y0=y0).cpu().numpy()
    ctxs = [D.Context(i) for i in range(a.gpus)]
    phase_ms, call_ms, out = [], [], None
    for it in range(a.steps + 1):
        t0 = time.perf_counter()
        out = ctxs[0].encode_sharded(px, a.gpus, 255, D.Options(), contexts=ctxs)
        call_ms.append((time.perf_counter() - t0) * 1e3)
        phase_ms.append(F.lib().dmmt_encode_sharded_last_ms())
    digest = hashlib.sha256(out).hexdigest()
    verified = "unchecked"
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "config5_sha256.json")))[str(n)]
        verified = "sha256 and length equal the oracle's file" if (gold["sha256"], gold["bytes"]) == (digest, len(out)) else "MISMATCH"
    except Exception:
        pass
    best = min(phase_ms[1:])
    print(json.dumps({"workload": f"one {n}x{n} 'smooth' image, dmmt_encode_sharded over {a.gpus} contexts in one process",
                      "phase_section_ms": best, "phase_section_ms_all": [round(x, 3) for x in phase_ms],
                      "mpixel_per_s": n * n / best / 1e3, "whole_call_ms": round(min(call_ms[1:]), 1),
                      "file_bytes": len(out), "verified": verified}))


if __name__ == "__main__":
    main()
