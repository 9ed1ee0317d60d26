"""BASELINE config 5: one synthetic 32768 x 32768 image sharded by MCU rows over the ranks of a torchrun
job (one process per GPU), the four small exchanges over NCCL (dmmt_jpeg_encoder_b200/sharded.py).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [--size 32768]
Each rank generates its own pixel rows on its device (not timed).  A step = the whole sharded encode
(K1 / exchange / histogram / all-reduce / tables / exchange / pack / exchange / stuff / gather to rank 0);
timed by wall clock between barriers + device synchronisation, max over ranks.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmmt_jpeg_encoder_b200 as D  # noqa: E402
from dmmt_jpeg_encoder_b200 import _ffi as F  # noqa: E402
from dmmt_jpeg_encoder_b200 import sharded as S  # noqa: E402
from dmmt_jpeg_encoder_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--phases", action="store_true", help="also report per-phase device and enqueue times (device exchange only)")
    ap.add_argument("--peer", action="store_true", help="no gather: K4 of every rank writes into rank 0's file over NVLink (CUDA IPC)")
    ap.add_argument("--host-exchange", action="store_true", help="exchange through host values (dmmt_shard_* phases)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = a.size
    opts = D.Options()
    rows = S.mcu_rows_total(n, opts)
    b, e = S.shard_rows(rows, world, rank)
    y0, y1 = S.pixel_row_range(n, opts, b, e)
    d_px = torch.empty((y1 - y0, n, 3), dtype=torch.uint8, device=dev)
    for s0 in range(y0, y1, 2048):
        s1 = min(y1, s0 + 2048)
        d_px[s0 - y0:s1 - y0] = synth.make("smooth", 5, s1 - s0, n, dev, y0=s0)
    torch.cuda.synchronize()
    ctx = D.Context(local, torch.cuda.current_stream().cuda_stream)
    be = S.CudaShardBackend(ctx, d_px.data_ptr(), n, n, F.FMT_U8, 255, opts, b, e)
    pf = S.PeerFile(be) if a.peer else None
    out = None
    times = []
    phase_dev, phase_cpu = {}, {}
    for it in range(a.warmup + a.steps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # the file stays in rank 0's HBM, like bench.py's `value`
        marks = []

        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev, time.perf_counter()))
        if a.phases and not a.host_exchange:
            mark("start")
        if a.host_exchange:
            out = S.encode_sharded(be, dev, to_host=False)
        elif a.peer:
            out = S.encode_sharded_peer(be, pf, to_host=False, mark=mark if a.phases else None)
        else:
            out = S.encode_sharded_device(be, to_host=False, mark=mark if a.phases else None)
        torch.cuda.synchronize()
        if marks and it >= a.warmup:
            for (_, e0, c0), (name, e1, c1) in zip(marks, marks[1:]):
                phase_dev[name] = phase_dev.get(name, 0.0) + e0.elapsed_time(e1) / a.steps
                phase_cpu[name] = phase_cpu.get(name, 0.0) + (c1 - c0) * 1e3 / a.steps
        dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it >= a.warmup:
            times.append(float(dt.item()))
    if a.phases and not a.host_exchange:
        allp = [None] * world
        dist.all_gather_object(allp, {"device_ms": phase_dev, "enqueue_ms": phase_cpu})
    if rank == 0:
        ms = sum(times) / len(times) * 1e3
        line = {"metric": "encoded MPixel/s", "config": {"workload": f"one synthetic {n}x{n} RGB u8 image ('smooth'), 4:2:0, "
                                                                     f"MCU-row shards over {world} GPU(s), NCCL exchanges, file gathered in rank 0's HBM"},
                "exchange": "host values" if a.host_exchange else ("device-resident, K4 writes into rank 0's file over NVLink (no gather)" if a.peer else "device-resident (one host sync)"), "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "value": n * n / ms / 1e3,
                "unit": "MPixel/s", "file_bytes": out.numel(), "bytes_per_pixel": out.numel() / (n * n),
                "timing": "wall clock between barriers, device synchronised, max over ranks"}
        if a.phases and not a.host_exchange:
            line["phases_per_rank"] = [{k: {n_: round(v, 3) for n_, v in d.items()} for k, d in p_.items()} for p_ in allp]
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if pf is not None:
        pf.close()
    be.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
