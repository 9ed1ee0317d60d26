"""Secondary measurements for the other BASELINE.json configs (not the headline bench):
  config 3: one 3840x2160 frame on one GPU (latency-bound: a single launch chain)
  config 5: one large square image on one GPU (the single-GPU reference point of the sharded config)
Device-resident, CUDA events on the plan's stream, W warm-ups then K timed encodes; prints one JSON line.
  python tools/bench_configs.py [--big 32768]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmmt_jpeg_encoder_b200 as D  # noqa: E402
from dmmt_jpeg_encoder_b200 import _ffi as F  # noqa: E402
from dmmt_jpeg_encoder_b200 import synth  # noqa: E402


def time_plan(ctx, stream, w, h, kind, steps, warmup, rows_at_a_time=4096):
    plan = D.Plan(ctx, w, h, F.FMT_U8, 255, D.Options(), 1)
    d_px = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    for y0 in range(0, h, rows_at_a_time):          # generate in slabs (int64 temporaries are 8x the image)
        y1 = min(h, y0 + rows_at_a_time)
        if kind == "photo":
            d_px[y0:y1] = synth.photo(0, h, w, "cuda")[y0:y1] if h <= rows_at_a_time else synth.smooth(0, y1 - y0, w, "cuda", y0)
        else:
            d_px[y0:y1] = synth.make(kind, 0, y1 - y0, w, "cuda", y0)
    d_out = torch.empty(plan.out_stride, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(warmup):
        plan.encode_device(d_px.data_ptr(), 1, d_out.data_ptr(), d_len.data_ptr())
    plan.status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            plan.encode_device(d_px.data_ptr(), 1, d_out.data_ptr(), d_len.data_ptr())
        e1.record(stream)
    torch.cuda.synchronize()
    plan.status()
    ms = e0.elapsed_time(e1) / steps
    plan.set_profiling(True)
    plan.encode_device(d_px.data_ptr(), 1, d_out.data_ptr(), d_len.data_ptr())
    plan.status()
    tm = plan.last_timings()
    n_bytes = int(d_len.item())
    plan.close()
    return {"width": w, "height": h, "kind": kind, "ms_per_image": ms, "mpixel_per_s": w * h / ms / 1e3,
            "file_bytes": n_bytes, "bytes_per_pixel": n_bytes / (w * h), "kernel_ms": tm}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", type=int, default=32768)
    a = ap.parse_args()
    stream = torch.cuda.Stream()
    ctx = D.Context(0, stream.cuda_stream)
    out = {"config3_single_4k": time_plan(ctx, stream, 3840, 2160, "photo", 50, 5),
           "config5_single_gpu_reference": time_plan(ctx, stream, a.big, a.big, "smooth", 5, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
