"""CPU simulator of one MCU-row shard (test infrastructure): implements the five phases of
dmmt_jpeg_encoder_b200.sharded.ShardBackend from the oracle's quantised coefficient stream with
plain Python loops (small images only), so the exchange arithmetic of encode_sharded can be
exercised over gloo without a GPU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


def _cat(v):
    a = abs(int(v))
    c = a.bit_length()
    bits = v if v > 0 else (1 << c) - 1 - a
    return c, bits


class SimShard:
    def __init__(self, px, preset, row_begin, row_end):
        r = O.encode(px, 255, preset, keep_planes=True)
        hr, vr = {0: (1, 1), 1: (2, 1), 2: (2, 2)}[preset]
        self.ypm, self.bpm = hr * vr, hr * vr + 2
        mcus_x = r.padded_width // (8 * hr)
        self.blocks = r.stream[row_begin * mcus_x * self.bpm: row_end * mcus_x * self.bpm]
        self.header = r.jpeg[: r.header_bytes]
        self.whole = r.jpeg

    def _comp(self, i):
        k = i % self.bpm
        return 0 if k < self.ypm else (1 if k == self.ypm else 2)

    def transform(self):
        last = [0, 0, 0]
        for i, b in enumerate(self.blocks):
            last[self._comp(i)] = int(b[0])
        return np.array(last, np.int16)

    def _tokens(self, seed):
        pred = [int(v) for v in seed]
        for i, b in enumerate(self.blocks):
            c = self._comp(i)
            t = 0 if c == 0 else 2
            cat, bits = _cat(int(b[0]) - pred[c])
            pred[c] = int(b[0])
            yield t, cat, bits, cat
            run = 0
            for v in b[1:]:
                v = int(v)
                if v == 0:
                    run += 1
                    continue
                while run > 15:
                    yield t + 1, 0xF0, 0, 0
                    run -= 16
                cat, bits = _cat(v)
                yield t + 1, (run << 4) | cat, bits, cat
                run = 0
            if run:
                yield t + 1, 0, 0, 0

    def histogram(self, seed_dc):
        self.seed = seed_dc
        h = np.zeros((4, 256), np.uint64)
        for t, sym, _, _ in self._tokens(seed_dc):
            h[t, sym] += 1
        return h.reshape(-1)

    def tables(self, global_hist):
        g = np.asarray(global_hist, np.uint64).reshape(4, 256)
        self.lut = []
        for t in range(4):
            sym, ln = O.build_table(g[t][: 256 if t & 1 else 16])
            code, cl = O.canonical_codes(sym, ln)
            self.lut.append((code, cl))
        self.bitstr = []
        for t, sym, bits, extra in self._tokens(self.seed):
            code, cl = self.lut[t]
            n = int(cl[sym])
            assert n > 0
            self.bitstr.append(format(int(code[sym]) >> (16 - n), f"0{n}b"))
            if extra:
                self.bitstr.append(format(bits, f"0{extra}b"))
        self.bitstr = "".join(self.bitstr)
        return len(self.bitstr)

    def pack(self, global_bit_offset, is_last):
        self.seed_bits = global_bit_offset & 7
        s = "0" * self.seed_bits + self.bitstr
        self.end = len(s)
        if is_last and len(s) % 8:
            s += "1" * (8 - len(s) % 8)
        s += "0" * (-len(s) % 8)
        self.bytes = bytearray(int(s[i:i + 8], 2) for i in range(0, len(s), 8))
        if is_last or self.end % 8 == 0:
            return 0, 0
        return self.bytes[self.end // 8], self.end % 8

    def stuff(self, prev_tail, prev_nbits, is_first, is_last):
        assert prev_nbits == self.seed_bits
        owned = (self.end + 7) // 8 if is_last else self.end // 8
        data = bytearray(self.bytes[:owned])
        if owned:
            data[0] |= prev_tail
        out = bytearray(self.header if is_first else b"")
        out += O.stuff_bytes(bytes(data))
        if is_last:
            out += b"\xff\xd9"
        return torch.frombuffer(bytearray(out), dtype=torch.uint8).clone() if out else torch.empty(0, dtype=torch.uint8)


class FailingShard(SimShard):
    """A shard whose `fail_phase` raises DmmtError(code) on rank `fail_rank` for the first `fail_times` attempts --
    the device-side failure of one shard (scan overflow, missing symbol ...) as the C ABI reports it."""

    def __init__(self, px, preset, row_begin, row_end, rank, fail_rank, fail_phase, code, fail_times):
        super().__init__(px, preset, row_begin, row_end)
        self.rank, self.fail_rank, self.fail_phase, self.code, self.left = rank, fail_rank, fail_phase, code, fail_times
        self.grown = 0

    def _maybe_fail(self, phase):
        from dmmt_jpeg_encoder_b200 import _ffi as F

        if self.rank == self.fail_rank and phase == self.fail_phase and self.left > 0:
            self.left -= 1
            raise F.DmmtError(self.code, f"simulated failure in {phase}")

    def transform(self):
        self._maybe_fail("transform")
        return super().transform()

    def histogram(self, seed_dc):
        self._maybe_fail("histogram")
        return super().histogram(seed_dc)

    def tables(self, global_hist):
        self._maybe_fail("tables")
        return super().tables(global_hist)

    def pack(self, global_bit_offset, is_last):
        self._maybe_fail("pack")
        return super().pack(global_bit_offset, is_last)

    def stuff(self, prev_tail, prev_nbits, is_first, is_last):
        self._maybe_fail("stuff")
        return super().stuff(prev_tail, prev_nbits, is_first, is_last)

    def grow(self):
        self.grown += 1


def failing_worker(rank, world, port, px, preset, q, fail_rank, fail_phase, code, fail_times):
    """every rank reports (rank, outcome): the file / None, or the code of the DmmtError it raised, + grow() calls"""
    import torch.distributed as dist

    from dmmt_jpeg_encoder_b200 import _ffi as F
    from dmmt_jpeg_encoder_b200 import sharded as S
    from dmmt_jpeg_encoder_b200.encoder import Options

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        opts = Options(preset, 8, 0)
        rows = S.mcu_rows_total(px.shape[0], opts)
        b, e = S.shard_rows(rows, world, rank)
        sim = FailingShard(px, preset, b, e, rank, fail_rank, fail_phase, code, fail_times)
        try:
            out = S.encode_sharded(sim, torch.device("cpu"))
            q.put((rank, "ok", out, sim.whole if rank == 0 else None, sim.grown))
        except F.DmmtError as err:
            q.put((rank, "error", err.code, None, sim.grown))
    finally:
        dist.destroy_process_group()


def worker(rank, world, port, px, preset, q):
    import torch.distributed as dist

    from dmmt_jpeg_encoder_b200 import sharded as S
    from dmmt_jpeg_encoder_b200.encoder import Options

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        opts = Options(preset, 8, 0)
        rows = S.mcu_rows_total(px.shape[0], opts)
        b, e = S.shard_rows(rows, world, rank)
        sim = SimShard(px, preset, b, e)
        out = S.encode_sharded(sim, torch.device("cpu"))
        if rank == 0:
            q.put((out, sim.whole))
    finally:
        dist.destroy_process_group()
