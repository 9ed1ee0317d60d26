"""torchrun worker (test infrastructure): one image, MCU-row shards, one process per GPU over NCCL.
Rank 0 compares the stitched file with the single-GPU encode and (small sizes) with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dmmt_jpeg_encoder_b200 as D  # noqa: E402
from dmmt_jpeg_encoder_b200 import _ffi as F  # noqa: E402
from dmmt_jpeg_encoder_b200 import sharded as S  # noqa: E402
from dmmt_jpeg_encoder_b200 import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = D.Context(local, torch.cuda.current_stream().cuda_stream)   # launches and NCCL share torch's stream
    failures = 0
    for (w, h, preset, use_oracle) in [(1000, 650, F.P420, True), (333, 97, F.P444, True), (640, 360, F.P422, True),
                                        (8192, 4096, F.P420, False)]:
        opts = D.Options(preset, 8, 0)
        rows = S.mcu_rows_total(h, opts)
        b, e = S.shard_rows(rows, world, rank)
        y0, y1 = S.pixel_row_range(h, opts, b, e)
        d_px = synth.make("smooth", 7, y1 - y0, w, dev, y0=y0)       # this rank's pixel rows only
        torch.cuda.synchronize()
        be = S.CudaShardBackend(ctx, d_px.data_ptr(), w, h, F.FMT_U8, 255, opts, b, e)
        assert be.pixel_bytes == d_px.numel(), (be.pixel_bytes, d_px.numel())
        out = S.encode_sharded(be, dev)
        out_dev = S.encode_sharded_device(be)                          # device-resident exchange: same bytes
        pf = S.PeerFile(be)                                            # K4 of every rank writes into rank 0's file
        out_peer = S.encode_sharded_peer(be, pf)
        out_peer2 = S.encode_sharded_peer(be, pf)                      # buffers and counters are reusable
        mb = S.PeerMailbox(be)                                         # the exchanges as the library's own kernels
        out_mb = S.encode_sharded_peer(be, pf, mailbox=mb)
        out_mb2 = S.encode_sharded_peer(be, pf, mailbox=mb)            # sequence numbers advance
        out_nccl_again = S.encode_sharded_peer(be, pf)                 # and the two forms can alternate
        mb.close()
        pf.close()
        if rank == 0:
            full = synth.make("smooth", 7, h, w, "cpu").numpy()
            whole = ctx.encode(full, 255, opts)
            ok = out == whole and out_dev == whole and out_peer == whole and out_peer2 == whole
            ok = ok and out_mb == whole and out_mb2 == whole and out_nccl_again == whole
            if use_oracle:
                from oracle import oracle as O

                ok = ok and out == O.encode(full, 255, preset).jpeg
            print(f"sharded {w}x{h} preset {preset} world {world}: {'OK' if ok else 'MISMATCH'} ({len(out)} bytes)", flush=True)
            failures += 0 if ok else 1
        be.close()
        dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
