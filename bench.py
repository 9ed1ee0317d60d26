#!/usr/bin/env python
"""bench.py -- encoded MPixel/s of the B200 encode path on BASELINE.json's batched workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): a batch of 1024
synthetic 1920x1080 RGB frames, 4:2:0, Annex-K tables; image i belongs to rank i mod N (round-robin,
no data-path collective).  A "step" = one pass of the whole encode path (K1..K5: colour/pad/
subsample/DCT/quantise -> histogram -> per-image Huffman tables -> bit pack -> byte stuffing ->
packed files) over the rank's share of the batch.

  value : whole-job MPixel/s (original pixels), inputs resident in HBM, outputs left in HBM,
          CUDA events on the launching stream, max over ranks.
  e2e   : the same through the host-buffer C-ABI call (dmmt_batch_encode_host): pinned host
          pixels -> H2D -> encode -> D2H of the files, every step.
  roofline / kernels : per-kernel CUDA-event times of one extra profiled pass (slots serialised),
          algorithmic bytes per launch / time vs MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline : the C oracle (restatement of the reference; the Rust crate cannot be built in
          this image) on the host cores, on a bounded sample of the same images.
  extra.config3 (N = 1): BASELINE config 3, ONE 3840x2160 frame: device microseconds per encode (events, chain
          replayed from a CUDA graph) and host-buffer microseconds per encode (dmmt_plan_encode_host_into).
  extra.config5 (N > 1): BASELINE config 5, ONE 32768x32768 image sharded by MCU rows over the N ranks, NCCL
          exchanges, every rank's K4 writing into rank 0's file over NVLink; verified against the ORACLE's
          SHA-256 of the same image (tests/golden/config5_sha256.json).

`--impl reference` times only that CPU arm (rank 0), with the same JSON shape.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "encoded MPixel/s"
UNIT = "MPixel/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=1024, help="global batch (images)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--kind", default="photo", choices=["photo", "grad", "uniform"])
    ap.add_argument("--sub-batch", type=int, default=256)
    ap.add_argument("--depth", type=int, default=2)
    ap.add_argument("--exact-sub-batch", action="store_true", help="use --sub-batch as given (no per-rank heuristic)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="images of the batch timed on the CPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra.config3 / extra.config5")
    ap.add_argument("--no-verify", action="store_true", help="kernel experiments only: do not compare with the oracle")
    ap.add_argument("--config5-size", type=int, default=32768)
    return ap.parse_args()


def hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (pynvml)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml as N

            N.nvmlInit()
            self.N = N
            self.h = N.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM)
        except Exception:
            self.N = None

    def _loop(self):
        N = self.N
        names = {
            getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM))
                try:
                    r = N.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.N is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._t:
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_arm(args, indices, steps, warmup):
    """Times the oracle (C restatement of the reference) on `indices` of the synthetic batch with
    one image per host thread (the reference is one image per process, src/lib.rs:59-77; a batch
    user runs one process per core).  Returns (MPixel/s, cores, sample description, ms per step)."""
    from concurrent.futures import ThreadPoolExecutor

    from dmmt_jpeg_encoder_b200 import synth
    from oracle import oracle as O

    O.build()
    cores = os.cpu_count() or 1
    imgs = [synth.make(args.kind, i, args.height, args.width).numpy() for i in indices]

    def one(px):
        return len(O.encode(px, 255, O.P420, 8, 0, nthreads=1).jpeg)

    def step():
        with ThreadPoolExecutor(cores) as ex:
            return sum(ex.map(one, imgs))

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    mpix = len(imgs) * args.width * args.height * steps / dt / 1e6
    # two more shapes of the same CPU code on ONE frame: a single core, and the reference's own shape --
    # one image per process, only the DCT fans out over the thread pool (lib.rs:62, transformer.rs:126-148)
    def one_frame(nthreads):
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            O.encode(imgs[0], 255, O.P420, 8, 0, nthreads=nthreads)
            best = min(best, time.perf_counter() - t)
        return args.width * args.height / best / 1e6
    cpu_arm.extra = {"one_core": {"value": one_frame(1), "unit": UNIT, "cores": 1, "sample": "one frame, best of 3"},
                     "reference_shaped": {"value": one_frame(cores), "unit": UNIT, "cores": cores,
                                          "sample": "one frame per process, DCT on all host threads in 700-block "
                                                    "jobs, everything else on one thread; best of 3"}}
    sample = (f"{len(imgs)} of the {args.images} {args.width}x{args.height} '{args.kind}' frames per step, "
              f"{steps} step(s), one frame per host thread ({cores} threads), C restatement of the reference "
              f"(oracle/, gcc -O2 -ffp-contract=off); timed region = JpegImageWriter::write_image equivalent")
    return mpix, cores, sample, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(args.cpu_sample, args.images)
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    mpix, cores, sample, ms = cpu_arm(args, list(range(n)), steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": f"synthetic ({args.kind}, int-hash generator, seeds 1234+i)",
        "config": {"workload": f"batch of {args.images} synthetic {args.width}x{args.height} RGB u8 frames -> baseline "
                               f"JPEG 4:2:0, Annex-K tables, per-image optimal Huffman tables (bounded CPU sample: "
                               f"{n} frames per step)"},
        "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         **getattr(cpu_arm, "extra", {})},
        "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is Rust-only and no Rust toolchain exists in this image: this arm is the C oracle "
                "(oracle/dmmt_oracle.c), a restatement of the reference's algorithm, not the Rust binary",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_cpus(index: int):
    """Restricts this process to the CPUs NVML reports as local to CUDA device `index` (looked up by PCI bus id: NVML and
    CUDA may enumerate differently), so that first-touch places the pinned host buffers of the e2e path on that GPU's
    NUMA node (one process per GPU).  Returns a description for the JSON line."""
    try:
        import pynvml as N
        import torch

        N.nvmlInit()
        pr = torch.cuda.get_device_properties(index)
        try:
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            h = N.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            bus, h = f"nvml index {index}", N.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 64
        words = N.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            lo, hi = min(cpus), max(cpus)
            return f"gpu {index} ({bus}): {len(cpus)} cpus {lo}-{hi}"
    except Exception as e:  # no NVML / not permitted: keep the inherited affinity
        return f"gpu {index}: unbound ({type(e).__name__})"
    return f"gpu {index}: unbound"


def copy_ceiling(torch, dev, h_in, n_bytes_in, h_out, n_bytes_out, steps):
    """Raw pinned-memory copy ceiling of this rank for the e2e traffic pattern: the step's H2D bytes and D2H bytes as
    plain cudaMemcpyAsync in 32 MB pieces on two streams (both directions in flight), nothing else.  Returns seconds
    per step; the caller turns the max over ranks into the job's copy-bound MPixel/s."""
    d_in = torch.empty(n_bytes_in, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(max(n_bytes_out, 1), dtype=torch.uint8, device=dev)
    t_in = torch.from_numpy(h_in.array)[:n_bytes_in]
    t_out = torch.from_numpy(h_out.array)[:max(n_bytes_out, 1)]
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    piece = 32 << 20

    def one():
        with torch.cuda.stream(s1):
            for o in range(0, n_bytes_in, piece):
                d_in[o:o + piece].copy_(t_in[o:o + piece], non_blocking=True)
        with torch.cuda.stream(s2):
            for o in range(0, n_bytes_out, piece):
                t_out[o:o + piece].copy_(d_out[o:o + piece], non_blocking=True)
    one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def run_config3(D, F, synth, torch, ctx, stream, dev):
    """BASELINE config 3: ONE 3840x2160 frame on one GPU.  device_us: CUDA events around 50 encodes of the
    device-resident chain (dmmt_plan_encode_device, replayed from a CUDA graph after the first two calls);
    host_us: wall clock per dmmt_plan_encode_host_into (pinned host pixels -> pinned host file, H2D + D2H inside)."""
    import numpy as np

    from oracle import oracle as O

    W, H = 3840, 2160
    plan = D.Plan(ctx, W, H, F.FMT_U8, 255, D.Options(), 1)
    d_px = synth.make("photo", 0, H, W, dev).contiguous()
    d_out = torch.empty(plan.out_stride, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    for _ in range(5):
        plan.encode_device(d_px.data_ptr(), 1, d_out.data_ptr(), d_len.data_ptr())
    plan.status()
    n_bytes = int(d_len.item())
    want = O.encode(d_px.cpu().numpy(), 255, O.P420).jpeg
    ok = d_out[:n_bytes].cpu().numpy().tobytes() == want
    K = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(K):
            plan.encode_device(d_px.data_ptr(), 1, d_out.data_ptr(), d_len.data_ptr())
        e1.record(stream)
    torch.cuda.synchronize()
    plan.status()
    device_us = e0.elapsed_time(e1) / K * 1e3
    launches = plan.last_launch_count()
    plan.set_profiling(True)
    plan.encode_device(d_px.data_ptr(), 1, d_out.data_ptr(), d_len.data_ptr())
    plan.status()
    tm = {k: round(v * 1e3, 2) for k, v in plan.last_timings().items()}
    plan.set_profiling(False)
    # host buffers
    h_in, h_out = D.PinnedBuffer(W * H * 3), D.PinnedBuffer(plan.out_stride)
    torch.from_numpy(h_in.array).view(H, W, 3).copy_(d_px)
    torch.cuda.synchronize()
    offs, lens = np.zeros(1, np.uint64), np.zeros(1, np.uint64)
    for _ in range(3):
        plan.encode_host_into(h_in.ptr, 1, h_out.ptr, plan.out_stride, offs, lens)
    ok = ok and h_out.array[: int(lens[0])].tobytes() == want
    t0 = time.perf_counter()
    for _ in range(20):
        plan.encode_host_into(h_in.ptr, 1, h_out.ptr, plan.out_stride, offs, lens)
    host_us = (time.perf_counter() - t0) / 20 * 1e6
    plan.close()
    h_in.close(), h_out.close()
    return {"workload": "one synthetic 3840x2160 RGB u8 frame ('photo') -> baseline JPEG 4:2:0, device-resident chain",
            "device_us": device_us, "mpixel_per_s": W * H / device_us, "host_us": host_us,
            "host_mpixel_per_s": W * H / host_us, "h2d_bytes": W * H * 3, "d2h_bytes": n_bytes + 16,
            "kernel_us": tm, "launches_per_encode": launches, "file_bytes": n_bytes,
            "verified": "byte-identical to the oracle (device-resident and host-buffer paths)" if ok else "MISMATCH"}


def run_bounds(D, F, synth, torch, ctx, stream, dev, W, H, n=128, steps=5, warmup=3):
    """the headline uses `photo`; `grad` (sparse: the reference's own dct_timing pattern) and `uniform` (iid
    noise: the worst case of the entropy stage) are reported as bounds.  Same device-resident chain on n frames."""
    from oracle import oracle as O

    out = {}
    for kind in ("grad", "uniform"):
        d_px = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
        for i in range(n):
            d_px[i] = synth.make(kind, i, H, W, dev)
        batch = D.Batch(ctx, W, H, F.FMT_U8, 255, D.Options(F.P420, 8, 0), n, 1)
        cap = n * W * H * 2
        d_dense = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        d_len = torch.zeros(n, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()

        def step():
            batch.encode_device(d_px.data_ptr(), n, d_dense.data_ptr(), cap, d_off.data_ptr(), d_len.data_ptr())
        grown = False
        step()
        try:
            batch.status()
        except F.DmmtError as e:                      # dense content: grow to the worst case once, like dmmt_encode does
            if e.code != F.E_OVERFLOW:
                raise
            batch.set_scan_capacity(batch.worst_case_scan_bytes())
            grown = True
        for _ in range(warmup):
            step()
        batch.status()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                step()
            e1.record(stream)
        torch.cuda.synchronize()
        batch.status()
        ms = e0.elapsed_time(e1) / steps
        lens, offs = d_len.cpu().numpy().astype("int64"), d_off.cpu().numpy().astype("int64")
        j = n - 1
        ok = d_dense[offs[j]: offs[j] + lens[j]].cpu().numpy().tobytes() == O.encode(d_px[j].cpu().numpy(), 255, O.P420).jpeg
        out[kind] = {"mpixel_per_s": n * W * H / ms / 1e3, "ms": ms, "images": n, "bytes_per_pixel_out": float(lens.sum()) / (n * W * H),
                     "scan_capacity": "worst case (after DMMT_E_OVERFLOW)" if grown else "default",
                     "verified": f"image {j} byte-identical to the oracle" if ok else "MISMATCH"}
        batch.close()
        del d_px, d_dense
        torch.cuda.empty_cache()
    return out


def run_config5(D, F, synth, torch, dist, local, dev, size, steps=5, warmup=2):
    """BASELINE config 5: ONE size x size image, MCU-row shards over the ranks (one process per GPU), the small
    exchanges by the library's own kernels over peer memory (and, for comparison, over NCCL on the same stream), every
    rank's K4 storing into rank 0's file over NVLink (CUDA IPC).
    Timed with CUDA events on every rank between barriers, max over ranks.  Rank 0 compares the SHA-256 of the file
    with the oracle's committed digest."""
    import hashlib

    from dmmt_jpeg_encoder_b200 import sharded as S

    rank, world = dist.get_rank(), dist.get_world_size()
    n = size
    opts = D.Options()
    rows = S.mcu_rows_total(n, opts)
    b, e = S.shard_rows(rows, world, rank)
    y0, y1 = S.pixel_row_range(n, opts, b, e)
    d_px = torch.empty((y1 - y0, n, 3), dtype=torch.uint8, device=dev)
    for s0 in range(y0, y1, 2048):
        s1 = min(y1, s0 + 2048)
        d_px[s0 - y0:s1 - y0] = synth.make("smooth", 5, s1 - s0, n, dev, y0=s0)
    torch.cuda.synchronize()
    ctx = D.Context(local, torch.cuda.current_stream().cuda_stream)   # launches and NCCL share torch's stream
    be = S.CudaShardBackend(ctx, d_px.data_ptr(), n, n, F.FMT_U8, 255, opts, b, e)
    pf = S.PeerFile(be)
    mb = S.PeerMailbox(be)

    def timed(mailbox):
        out, ms, phases = None, [], {}
        for it in range(warmup + steps):
            marks = []

            def mark(name):
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))
            dist.barrier()
            torch.cuda.synchronize()
            mark("start")
            out = S.encode_sharded_peer(be, pf, to_host=False, mark=mark, mailbox=mailbox)
            mark("end")
            torch.cuda.synchronize()
            t = torch.tensor([marks[0][1].elapsed_time(marks[-1][1])], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it >= warmup:
                ms.append(float(t.item()))
                for (_, a0), (name, a1) in zip(marks, marks[1:]):
                    phases[name] = phases.get(name, 0.0) + a0.elapsed_time(a1) / steps
        digest, nbytes = None, 0
        if rank == 0:
            digest, nbytes = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest(), out.numel()
        return sum(ms) / len(ms), {k: round(v, 3) for k, v in phases.items()}, digest, nbytes

    t_nccl, ph_nccl, dg_nccl, nb_nccl = timed(None)
    t, phases, digest, nbytes = timed(mb)
    res = None
    if rank == 0:
        verified = "unchecked (no committed oracle digest for this size)"
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "config5_sha256.json")))[str(n)]
            same = (gold["sha256"], gold["bytes"]) == (digest, nbytes) == (dg_nccl, nb_nccl)
            verified = ("sha256 and length equal the oracle's file (tests/golden/config5_sha256.json), both exchange forms"
                        if same else "MISMATCH with the oracle's digest")
        except Exception:
            pass
        res = {"workload": f"one synthetic {n}x{n} RGB u8 image ('smooth') -> baseline JPEG 4:2:0, MCU-row shards over {world} "
                           f"ranks; last DCs / histograms / bit counts / tails / byte counts exchanged by the library's own "
                           f"kernels over peer memory (mailboxes mapped through CUDA IPC, no collective library on the data "
                           f"path); K4 of every rank writes into rank 0's file over NVLink",
               "ms": t, "mpixel_per_s": n * n / t / 1e3,
               "n_gpus": world, "steps": steps, "warmup": warmup, "file_bytes": nbytes,
               "phases_ms_rank0": phases,
               "nccl_exchanges": {"ms": t_nccl, "mpixel_per_s": n * n / t_nccl / 1e3, "phases_ms_rank0": ph_nccl,
                                  "what": "the same phases with NCCL all-gathers / all-reduce on the same stream"},
               "timing": "CUDA events on each rank's stream between barriers, max over ranks", "verified": verified}
    mb.close()
    pf.close()
    be.close()
    ctx.close()
    del d_px
    torch.cuda.empty_cache()
    return res


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import dmmt_jpeg_encoder_b200 as D
    from dmmt_jpeg_encoder_b200 import _ffi as F
    from dmmt_jpeg_encoder_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL prints its version there)
    # are routed to stderr until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available() or F.lib().dmmt_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the encode path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_cpus(local)                         # pinned staging buffers land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    W, H, K, WU = args.width, args.height, args.steps, max(args.warmup, 0)
    mine = list(range(rank, args.images, world))           # round-robin: image i -> rank i mod N
    n = len(mine)
    img_bytes = W * H * 3
    pW, pH = (W + 15) // 16 * 16, (H + 15) // 16 * 16

    # ---- synthetic inputs, generated on the device (not timed)
    d_px = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
    for j, i in enumerate(mine):
        d_px[j] = synth.make(args.kind, i, H, W, dev)
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    ctx = D.Context(local, stream.cuda_stream)
    dev_sub = max(1, min(args.sub_batch, max(16, (n + 1) // 2)))   # at least 2 sub-batches so the streams overlap
    if args.exact_sub_batch:
        dev_sub = max(1, min(args.sub_batch, n))
    batch = D.Batch(ctx, W, H, F.FMT_U8, 255, D.Options(F.P420, 8, 0), dev_sub, args.depth)
    dense_cap = n * (img_bytes // 4)
    d_dense = torch.empty(dense_cap, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step():
        batch.encode_device(d_px.data_ptr(), n, d_dense.data_ptr(), dense_cap, d_off.data_ptr(), d_len.data_ptr())

    for _ in range(max(WU, 1)):
        step()
    batch.status()                                          # raises on any device-side error
    lens = d_len.cpu().numpy().astype(np.int64)
    offs = d_off.cpu().numpy().astype(np.int64)
    file_bytes = int(lens.sum())

    # ---- the checker (outside every timed region): image 0 of this rank byte-identical to the oracle
    verified = None
    if rank == 0 and args.no_verify:
        verified, want = "NOT VERIFIED (--no-verify)", None
    elif rank == 0:
        from oracle import oracle as O

        check = sorted({0, 1, n // 2, n - 1})                 # first, second, middle and last image of this rank's share
        for j in check:
            got = d_dense[offs[j]: offs[j] + lens[j]].cpu().numpy().tobytes()
            w_j = O.encode(d_px[j].cpu().numpy(), 255, O.P420).jpeg
            if j == 0:
                want = w_j
            if got != w_j:
                raise SystemExit(f"bench: CUDA output of image {j} differs from the oracle -- numbers would be meaningless")
        verified = f"images {check} of rank 0 byte-identical to the oracle"

    # ---- value: device-resident, CUDA events on the context's stream
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(K):
            step()
        ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    batch.status()
    ms_total = reduce(ev0.elapsed_time(ev1), dist.ReduceOp.MAX if world > 1 else None)
    launches = reduce(float(batch.last_launch_count() * K), dist.ReduceOp.SUM if world > 1 else None)
    total_px = args.images * W * H
    value = total_px * K / (ms_total / 1e3) / 1e6

    # ---- e2e: host buffers through dmmt_batch_encode_host, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        h_in = D.PinnedBuffer(n * img_bytes)
        h_out_cap = max(file_bytes * 2, 1 << 20)
        h_out = D.PinnedBuffer(h_out_cap)
        torch.from_numpy(h_in.array).view(n, H, W, 3).copy_(d_px)      # stage the inputs on the host once
        torch.cuda.synchronize()
        h_offs, h_lens = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        # sub-batches of >= 64 frames (398 MB of pixels) keep every copy long; 4 slots = up to 3 H2D copies queued behind
        # the running one, so the copy engine never waits for the host
        e2e_sub = max(1, min(args.sub_batch, max(min(64, n), n // 8)))
        e2e_depth = max(args.depth, 4)
        batch_h = D.Batch(ctx, W, H, F.FMT_U8, 255, D.Options(F.P420, 8, 0), e2e_sub, e2e_depth)

        def estep():
            batch_h.encode_host(h_in.ptr, n, h_out.ptr, h_out_cap, h_offs, h_lens)

        for _ in range(max(min(WU, 2), 1)):
            estep()
        assert int(h_lens.sum()) == file_bytes
        if rank == 0 and want is not None:
            assert h_out.array[int(h_offs[0]): int(h_offs[0] + h_lens[0])].tobytes() == want
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            estep()                                         # synchronous at return
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        dt = reduce(dt, dist.ReduceOp.MAX if world > 1 else None)
        d2h = reduce(float(int(h_offs[-1] + (h_lens[-1] + 15) // 16 * 16) + 16 * n + 8 * (n + (n + e2e_sub - 1) // e2e_sub)),
                     dist.ReduceOp.SUM if world > 1 else None)
        h2d_total = reduce(float(n * img_bytes), dist.ReduceOp.SUM if world > 1 else None)
        # the raw copy ceiling of the same traffic on the same buffers, all ranks at once
        barrier()
        ct = copy_ceiling(torch, dev, h_in, n * img_bytes, h_out, file_bytes, max(2, min(K, 5)))
        barrier()
        ct = reduce(ct, dist.ReduceOp.MAX if world > 1 else None)
        bindings = [numa]
        if world > 1:
            bindings = [None] * world
            dist.all_gather_object(bindings, numa)
        e2e = {"value": total_px * K / dt / 1e6, "unit": UNIT, "ms_per_step": dt / K * 1e3,
               "h2d_bytes_per_step": int(h2d_total), "d2h_bytes_per_step": int(d2h),
               "api": "dmmt_batch_encode_host (pinned host pixels -> packed host files), all ranks",
               "sub_batch": e2e_sub, "streams": e2e_depth, "host_binding": bindings,
               "h2d_gbs": h2d_total * K / dt / 1e9,
               "copy_ceiling": {"value": total_px / ct / 1e6, "unit": UNIT, "ms_per_step": ct * 1e3,
                                "h2d_gbs": h2d_total / ct / 1e9,
                                "what": "the same H2D + D2H bytes as plain cudaMemcpyAsync from / to the same pinned buffers, "
                                        "all ranks at once, no kernels"},
               "frac_of_copy_ceiling": (total_px * K / dt) / (total_px / ct)}
        batch_h.close()
        h_in.close(), h_out.close()

    # ---- per-kernel pass (slots serialised, events between kernels) -> roofline
    peak, peak_src = hbm_peak()
    batch.set_profiling(True)
    step()
    batch.status()
    tm = batch.last_timings()
    batch.set_profiling(False)
    n_sub = (n + batch.sub_batch - 1) // batch.sub_batch
    coef_bytes = 2 * (pW * pH * 3 // 2)                     # i16 coefficients per image, 4:2:0
    scan_b = (file_bytes / n) - 330.0                       # ~ stuffed scan bytes per image
    n_blocks = (pW // 8) * (pH // 8) * 3 // 2
    fused = batch.uses_fused_path()
    # algorithmic (compulsory) bytes per IMAGE, BASELINE.md section 3.  On the fused 4:2:0 path K1 also
    # does K2's work on-chip (the coefficient stream is never written or re-read), so the fused kernel is
    # credited with the per-stage figures of BOTH stages it replaces: 3 + 3 (K1) + 3 (K2) B/px.
    alg = {
        "k1_transform": img_bytes + coef_bytes + (coef_bytes if fused else 0),
        "k2_histogram": 0 if fused else coef_bytes,
        "k3_pack": coef_bytes + scan_b + 4 * n_blocks / 256.0,
        "k4_stuff": 2 * scan_b,
        "k5_compact": 2 * (file_bytes / n),
    }
    kernels = {}
    for name, per_img in alg.items():
        ms = tm[name]
        if ms <= 0 or per_img <= 0:
            continue
        gbs = per_img * n / (ms / 1e3) / 1e9
        kernels[name] = {"ms_per_step": ms, "launches_per_step": n_sub * (2 if name in ("k3_pack", "k5_compact") else 1),
                         "avg_launch_ms": ms / n_sub, "algorithmic_bytes_per_launch": per_img * n / n_sub,
                         "achieved_gbs": gbs, "frac": gbs / peak}
    kernels["k2b_tables"] = {"ms_per_step": tm["k2b_tables"], "launches_per_step": 2 * n_sub}
    if fused:
        kernels["k2_histogram"] = {"ms_per_step": tm["k2_histogram"], "launches_per_step": n_sub,
                                   "note": "fused path: only k2_fix_dc (tile-boundary DC tokens) runs here"}
        kernels["k1_transform"]["note"] = ("fused transform + tokenise + histogram kernel (k1_transform_p420<FMT, true>); "
                                           "algorithmic bytes = K1 6 B/px + K2 3 B/px (padded)")
        kernels["k1_transform"]["frac_at_6_bytes_per_px"] = (img_bytes + coef_bytes) * n / (tm["k1_transform"] / 1e3) / 1e9 / peak
    dom = max((k for k in kernels if "achieved_gbs" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    traffic, traffic_src = None, None
    try:  # dram__bytes_read + write of one ncu --set full capture, rescaled to this run's images per launch
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
        if t:
            traffic = t["dram_bytes_per_launch"] / t["images_per_launch"] * (n / n_sub)
            traffic_src = f"ncu --set full capture of {t.get('from', 'profiles/')}, rescaled to {n // n_sub} images per launch"
    except Exception:
        pass
    # what the dominant kernel really moves, measured in THIS run: the pixels it reads plus the tokens it wrote (counted
    # on the device: dmmt_plan_fetch(DMMT_FETCH_TOKEN_COUNT) on the first images of this rank's share) plus the per-tile
    # records (token count, last DCs, DC token positions) -- the coefficient stream of the 6 / 9 B/px figures never exists
    live = None
    if fused and dom == "k1_transform":
        m = min(n, 16)
        probe = D.Plan(ctx, W, H, F.FMT_U8, 255, D.Options(F.P420, 8, 0), m)
        d_o = torch.empty(m * probe.out_stride, dtype=torch.uint8, device=dev)
        probe.encode_device(d_px.data_ptr(), m, d_o.data_ptr())
        probe.status()
        tokens = sum(probe.fetch(F.FETCH_TOKEN_COUNT, i) for i in range(m)) / m
        probe.close()
        del d_o
        tiles = ((pW + 255) // 256) * (pH // 16)
        live = {"tokens_per_image": tokens, "bytes_per_image": img_bytes + 4 * tokens + tiles * (4 + 8 + 8),
                "sample_images": m}
        live["bytes_per_px"] = live["bytes_per_image"] / (W * H)
        live["gbs"] = live["bytes_per_image"] * n / (tm["k1_transform"] / 1e3) / 1e9
    k1 = kernels.get("k1_transform", {})
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"],
                "frac_k1_6Bpx": k1.get("frac_at_6_bytes_per_px", k1.get("frac")),
                "frac_dram": live["gbs"] / peak if live else None,
                "accounting": {"frac": "SURVEY 8d fused accounting: K1's 6 B/px + K2's 3 B/px (the stage the kernel absorbed)",
                               "frac_k1_6Bpx": "3 B/px read + 3 B/px of i16 coefficients, as if the coefficient stream were written",
                               "frac_dram": "bytes the kernel really moves (pixels + tokens + per-tile records, measured in "
                                            "this run) / time / peak"},
                "dram_live": live, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "share_of_step": kernels[dom]["ms_per_step"] / tm["total"] if tm["total"] > 0 else None,
                "k1_frac": k1.get("frac"), "fused_k1_k2": fused}

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m, cores, sample, _ = cpu_arm(args, mine[: min(args.cpu_sample, n)], 2, 1)
        cpu = {"value": m, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               **getattr(cpu_arm, "extra", {})}

    # ---- the other BASELINE configs, driver-visible: config 3 at N = 1, config 5 at N > 1
    extra = {}
    if not args.no_extra:
        batch.close()
        del d_dense
        torch.cuda.empty_cache()
        if world == 1:
            try:
                extra["config3"] = run_config3(D, F, synth, torch, ctx, stream, dev)
            except Exception as e:  # the headline line must survive a failure of the side measurement
                extra["config3"] = {"error": f"{type(e).__name__}: {e}"}
            del d_px
            torch.cuda.empty_cache()
            try:
                extra["bounds"] = run_bounds(D, F, synth, torch, ctx, stream, dev, W, H)
            except Exception as e:
                extra["bounds"] = {"error": f"{type(e).__name__}: {e}"}
        else:
            del d_px
            torch.cuda.empty_cache()
            try:
                extra["config5"] = run_config5(D, F, synth, torch, dist, local, dev, args.config5_size)
            except Exception as e:
                extra["config5"] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": WU,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": f"synthetic ({args.kind}, int-hash generator, seeds 1234+i, generated on device)",
            "config": {"workload": f"batch of {args.images} synthetic {W}x{H} RGB u8 frames -> baseline JPEG 4:2:0, "
                                   f"Annex-K tables, per-image optimal Huffman tables; image i on rank i mod N",
                       "images_per_rank": n, "sub_batch": batch.sub_batch, "streams": batch.depth,
                       "l2": f"inputs larger than L2 ({n * img_bytes / 1e6:.0f} MB of pixels per rank per step, no reuse)",
                       "bytes_per_pixel_out": file_bytes / (n * W * H), "verified": verified},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "kernels": kernels, "kernel_pass_total_ms": tm["total"],
            "cpu_baseline": cpu, "extra": extra,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    batch.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        rc = 1
        for attempt in range(3):   # the free port can be taken again before torchrun binds it: try another one
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.abspath(__file__)] + sys.argv[1:]
            p = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
            sys.stderr.write(p.stderr)
            rc = p.returncode
            if rc == 0 or "EADDRINUSE" not in p.stderr:
                break
        raise SystemExit(rc)
    run_ours(args)


if __name__ == "__main__":
    main()
