"""One image, MCU-row shards, one process per GPU (BASELINE config 5).

Rank r owns MCU rows [r*M/N, (r+1)*M/N).  The kernels run through the dmmt_shard_* C ABI; the four
small exchanges run over torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests):

  1. all_gather of the last quantised DC of (Y, Cb, Cr)     -> DC predictor seeds
  2. all_reduce(sum) of the 4 symbol histograms (1024 x i64) -> image-global Huffman tables
  3. all_gather of the entropy-coded bit counts + tail bits  -> global bit offsets
  4. all_gather of the stuffed byte counts                   -> byte offsets; send/recv to rank 0

`ShardBackend` is the seam: `CudaShardBackend` is the product; the CPU tests plug in a simulator so
the exchange arithmetic is covered without a GPU.  On one NVLink node the peer path (`encode_sharded_peer`) needs no
collective library at all: `PeerFile` lets K4 of every rank store into the destination rank's file, `PeerMailbox` moves
the exchanged values with the library's own kernels (dmmt_shard_launch_post / dmmt_shard_launch_collect).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as _dist

from . import _ffi as F
from .encoder import Context, Options


class _Collectives:
    """torch.distributed with the library's error model: a collective that fails (NCCL / gloo error, a rank that went
    away, a time-out) surfaces as DmmtError(DMMT_E_NCCL), the code include/dmmt_cuda.h reserves for the exchanges of
    the sharded path, instead of a backend-specific exception."""

    def __getattr__(self, name):
        fn = getattr(_dist, name)
        if not callable(fn) or isinstance(fn, type) or name in ("get_rank", "get_world_size", "get_global_rank", "isend", "irecv"):
            return fn

        def call(*a, **kw):
            try:
                return fn(*a, **kw)
            except F.DmmtError:
                raise
            except Exception as e:  # noqa: BLE001 -- every backend raises its own type
                raise F.DmmtError(F.E_NCCL, f"torch.distributed.{name}: {type(e).__name__}: {e}") from e
        return call


dist = _Collectives()


def shard_rows(total_mcu_rows: int, world: int, rank: int) -> tuple[int, int]:
    return rank * total_mcu_rows // world, (rank + 1) * total_mcu_rows // world


def mcu_rows_total(height: int, options: Options) -> int:
    vr = 2 if options.subsampling == F.P420 else 1
    return (height + 8 * vr - 1) // (8 * vr)


def pixel_row_range(height: int, options: Options, row_begin: int, row_end: int) -> tuple[int, int]:
    vr = 2 if options.subsampling == F.P420 else 1
    return row_begin * 8 * vr, min(height, row_end * 8 * vr)


class ShardBackend:
    """The five local phases of one shard."""

    def transform(self) -> np.ndarray: ...                                   # -> last_dc i16[3]
    def histogram(self, seed_dc: np.ndarray) -> np.ndarray: ...              # -> u64[1024]
    def tables(self, global_hist: np.ndarray) -> int: ...                    # -> local bits
    def pack(self, global_bit_offset: int, is_last: bool) -> tuple[int, int]: ...  # -> tail byte, nbits
    def stuff(self, prev_tail: int, prev_nbits: int, is_first: bool, is_last: bool) -> torch.Tensor: ...
    def grow(self) -> None: ...                                              # worst-case capacity after an overflow


class _DevPtr:
    """Zero-copy view of device memory owned by the C library as a torch tensor."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class CudaShardBackend(ShardBackend):
    def __init__(self, ctx: Context, d_pixels: int, full_width: int, full_height: int, fmt: int, max_value: int,
                 options: Options, row_begin: int, row_end: int):
        self.ctx = ctx
        self._h = C.c_void_p()
        o = options.c()
        F.check(F.lib().dmmt_shard_create(ctx.handle, full_width, full_height, fmt, max_value, C.byref(o),
                                          row_begin, row_end, C.byref(self._h)), "dmmt_shard_create")
        self.d_pixels = d_pixels

    @property
    def pixel_bytes(self) -> int:
        return F.lib().dmmt_shard_pixel_bytes(self._h)

    @property
    def pixel_offset(self) -> int:
        return F.lib().dmmt_shard_pixel_offset(self._h)

    def launch_count(self) -> int:
        return F.lib().dmmt_shard_launch_count(self._h)

    def transform(self):
        out = (C.c_int16 * 3)()
        F.check(F.lib().dmmt_shard_transform(self._h, C.c_void_p(self.d_pixels), out), "dmmt_shard_transform")
        return np.array(out[:], np.int16)

    def histogram(self, seed_dc):
        seed = (C.c_int16 * 3)(*[int(v) for v in seed_dc])
        h = np.zeros(1024, np.uint64)
        F.check(F.lib().dmmt_shard_histogram(self._h, seed, h.ctypes.data_as(F._U64P)), "dmmt_shard_histogram")
        return h

    def tables(self, global_hist):
        g = np.ascontiguousarray(global_hist, np.uint64)
        bits = C.c_uint64()
        F.check(F.lib().dmmt_shard_tables(self._h, g.ctypes.data_as(F._U64P), C.byref(bits)), "dmmt_shard_tables")
        return bits.value

    def pack(self, global_bit_offset, is_last):
        tb, tn = C.c_uint8(), C.c_int()
        F.check(F.lib().dmmt_shard_pack(self._h, global_bit_offset, int(is_last), C.byref(tb), C.byref(tn)),
                "dmmt_shard_pack")
        return tb.value, tn.value

    def stuff(self, prev_tail, prev_nbits, is_first, is_last):
        p, n = C.c_void_p(), C.c_uint64()
        F.check(F.lib().dmmt_shard_stuff(self._h, prev_tail, prev_nbits, int(is_first), int(is_last),
                                         C.byref(p), C.byref(n)), "dmmt_shard_stuff")
        if n.value == 0:
            return torch.empty(0, dtype=torch.uint8, device=f"cuda:{self.ctx.device}")
        return torch.as_tensor(_DevPtr(p.value, n.value), device=f"cuda:{self.ctx.device}")

    # ---- device-resident exchange (asynchronous launches, device pointers) ----
    def launch_transform(self, d_last_dc4: int):
        F.check(F.lib().dmmt_shard_launch_transform(self._h, C.c_void_p(self.d_pixels), C.c_void_p(d_last_dc4)),
                "dmmt_shard_launch_transform")

    def launch_histogram(self, d_seed_dc4: int, d_hist: int):
        F.check(F.lib().dmmt_shard_launch_histogram(self._h, C.c_void_p(d_seed_dc4) if d_seed_dc4 else None,
                                                    C.c_void_p(d_hist)), "dmmt_shard_launch_histogram")

    def launch_tables(self, d_global_hist: int, d_local_bits: int):
        F.check(F.lib().dmmt_shard_launch_tables(self._h, C.c_void_p(d_global_hist), C.c_void_p(d_local_bits)),
                "dmmt_shard_launch_tables")

    def launch_pack(self, d_bit_offset: int, is_last: bool, d_tail2: int):
        F.check(F.lib().dmmt_shard_launch_pack(self._h, C.c_void_p(d_bit_offset), int(is_last), C.c_void_p(d_tail2)),
                "dmmt_shard_launch_pack")

    def launch_stuff(self, d_all_tail2: int, d_all_offs: int, d_all_bits: int, rank: int, world: int, d_n_bytes: int) -> int:
        p = C.c_void_p()
        F.check(F.lib().dmmt_shard_launch_stuff(self._h, C.c_void_p(d_all_tail2), C.c_void_p(d_all_offs),
                                                C.c_void_p(d_all_bits), rank, world, C.byref(p), C.c_void_p(d_n_bytes)),
                "dmmt_shard_launch_stuff")
        return p.value

    def launch_count_bytes(self, d_all_tail2: int, d_all_offs: int, d_all_bits: int, rank: int, world: int, d_n_bytes: int):
        F.check(F.lib().dmmt_shard_launch_count_bytes(self._h, C.c_void_p(d_all_tail2), C.c_void_p(d_all_offs),
                                                      C.c_void_p(d_all_bits), rank, world, C.c_void_p(d_n_bytes)),
                "dmmt_shard_launch_count_bytes")

    def launch_stuff_into(self, d_all_offs: int, rank: int, world: int, d_file: int, capacity: int, d_byte_offset: int,
                          d_result2: int):
        F.check(F.lib().dmmt_shard_launch_stuff_into(self._h, C.c_void_p(d_all_offs), rank, world, C.c_void_p(d_file),
                                                     capacity, C.c_void_p(d_byte_offset), C.c_void_p(d_result2)),
                "dmmt_shard_launch_stuff_into")

    def launch_error(self, d_err: int):
        F.check(F.lib().dmmt_shard_launch_error(self._h, C.c_void_p(d_err)), "dmmt_shard_launch_error")

    def grow(self):
        """worst-case scan capacity (after DMMT_E_OVERFLOW on ANY shard of the image; every rank calls it)"""
        F.check(F.lib().dmmt_shard_set_scan_capacity(self._h, F.lib().dmmt_shard_worst_case_scan_bytes(self._h)),
                "dmmt_shard_set_scan_capacity")
        self._xbuf = self._pbuf = None

    @property
    def out_stride(self) -> int:
        return F.lib().dmmt_shard_out_stride(self._h)

    def status(self):
        F.check(F.lib().dmmt_shard_status(self._h), "dmmt_shard_status")

    def close(self):
        if self._h:
            F.lib().dmmt_shard_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Phases:
    """Runs the local phases of one rank so that a failure never strands the other ranks in a collective: the error
    code of a failed phase rides on the NEXT exchange (an extra slot of its payload), every rank sees the codes of
    all ranks and raises the same DmmtError -- or, for DMMT_E_OVERFLOW, grows its shard and runs again."""

    def __init__(self):
        self.err = 0

    def run(self, fn, *args, default=None):
        if self.err:
            return default                      # already failed: keep going through the collectives with dummies
        try:
            return fn(*args)
        except F.DmmtError as e:
            self.err = e.code
            return default


def _raise_first(codes, where):
    for c in codes:
        if c:
            raise F.DmmtError(int(c), where)


def encode_sharded(backend: ShardBackend, device: torch.device, group=None, dst: int = 0,
                   timings: dict | None = None, to_host: bool = True):
    """Runs the phases of this rank's shard with the four exchanges; returns the file on `dst`
    (bytes, or the uint8 tensor on `device` when to_host is False) and None on the other ranks.
    A device-side failure of any shard raises the same DmmtError on EVERY rank; an overflow of the default scan
    capacity makes every rank grow its shard to the worst case and run the phases once more."""
    for attempt in range(2):
        try:
            return _encode_sharded_once(backend, device, group, dst, timings, to_host)
        except F.DmmtError as e:
            if e.code != F.E_OVERFLOW or attempt:
                raise
            backend.grow()                       # every rank arrives here: the code was all-gathered


def _encode_sharded_once(backend, device, group, dst, timings, to_host):
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    is_first, is_last = rank == 0, rank == world - 1
    ph = _Phases()

    def gather(values):
        """all-gather of this rank's int64 values + its error code; raises on every rank if any rank failed"""
        t = torch.tensor(list(values) + [ph.err], dtype=torch.int64, device=device)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        rows = [x.cpu().tolist() for x in out]
        _raise_first([r[-1] for r in rows], "sharded encode")
        return [r[:-1] for r in rows]

    # phase 1 + exchange 1: DC predictor seeds (categorize.rs:157-161 never resets the chain)
    last_dc = ph.run(backend.transform, default=np.zeros(3, np.int16))
    all_dc = gather([int(v) for v in last_dc])
    seed = np.array(all_dc[rank - 1], np.int16) if rank else np.zeros(3, np.int16)

    # phase 2 + exchange 2: image-global histograms (transformer.rs:201-217); the error slot is summed too
    h = ph.run(backend.histogram, seed, default=np.zeros(1024, np.uint64))
    hist = torch.cat([torch.from_numpy(np.asarray(h).astype(np.int64)), torch.tensor([1 if ph.err else 0])]).to(device)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    failed = int(hist[-1].item()) != 0
    if failed:
        gather([])                               # somebody failed: this exchange tells every rank the code and raises

    # phase 3 + exchange 3: bit offsets
    bits = ph.run(backend.tables, hist[:-1].cpu().numpy().astype(np.uint64), default=0)
    all_bits = [r[0] for r in gather([bits])]
    bit_off = [0]
    for b in all_bits:
        bit_off.append(bit_off[-1] + b)

    # phase 4 + exchange of the trailing partial bytes
    tail_byte, tail_nbits = ph.run(backend.pack, bit_off[rank], is_last, default=(0, 0))
    tails = [r[0] for r in gather([tail_byte, tail_nbits])]
    # a shard that does not complete a byte hands its predecessor's bits on
    for r in range(1, world - 1):
        if (bit_off[r] & 7) + all_bits[r] < 8:
            tails[r] |= tails[r - 1]
    prev_tail = tails[rank - 1] if rank else 0
    prev_nbits = bit_off[rank] & 7

    # phase 5 + exchange 4: byte counts, then the bytes travel to dst
    mine = ph.run(backend.stuff, prev_tail, prev_nbits, is_first, is_last, default=torch.empty(0, dtype=torch.uint8))
    all_n = [r[0] for r in gather([mine.numel()])]
    if timings is not None:
        timings["bytes"] = all_n
        timings["bits"] = all_bits
    if rank == dst:
        out = torch.empty(sum(all_n), dtype=torch.uint8, device=device)
        off = 0
        for r in range(world):
            if all_n[r]:
                if r == rank:
                    out[off:off + all_n[r]].copy_(mine)
                else:
                    dist.recv(out[off:off + all_n[r]], src=r, group=group)
            off += all_n[r]
        return out.cpu().numpy().tobytes() if to_host else out
    if mine.numel():
        dist.send(mine.contiguous(), dst=dst, group=group)
    return None


def _retry_on_overflow(fn, backend, *args, **kw):
    """fn raises the same DmmtError on every rank (the codes were all-gathered); on DMMT_E_OVERFLOW every rank grows
    its shard to the worst-case capacity (and the peer file with it) and the phases run once more"""
    for attempt in range(2):
        try:
            return fn(backend, *args, **kw)
        except F.DmmtError as e:
            if e.code != F.E_OVERFLOW or attempt:
                raise
            backend.grow()
            for a in args:
                if isinstance(a, PeerFile):
                    a.regrow(backend)


def encode_sharded_device(backend: CudaShardBackend, group=None, dst: int = 0, to_host: bool = True, mark=None):
    """Same result as encode_sharded, but every exchanged value stays in device memory: the five phases are
    asynchronous launches on the context's stream (which must be torch's current stream) and the collectives
    are NCCL calls on device tensors, so there is ONE host synchronisation in the whole encode (the byte
    counts, needed to size the final send / recv).  The small exchange tensors are cached on the backend.
    Every rank's device error flag is all-gathered with the byte counts: a failed shard raises on every rank
    (overflow: every rank grows its shard and the phases run once more)."""
    return _retry_on_overflow(_encode_sharded_device_once, backend, group, dst, to_host, mark)


def _encode_sharded_device_once(backend: CudaShardBackend, group=None, dst: int = 0, to_host: bool = True, mark=None):
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", backend.ctx.device)
    buf = getattr(backend, "_xbuf", None)
    if buf is None or buf["world"] != world:
        i32, i64 = dict(dtype=torch.int32, device=dev), dict(dtype=torch.int64, device=dev)
        buf = backend._xbuf = {"world": world, "last": torch.empty(4, **i32), "all_dc": torch.empty(4 * world, **i32),
                               "hist": torch.empty(1024, **i64), "bits": torch.empty(1, **i64),
                               "all_bits": torch.empty(world, **i64), "offs": torch.empty(world, **i64),
                               "tail": torch.empty(2, **i32), "all_tail": torch.empty(2 * world, **i32),
                               "n_bytes": torch.zeros(2, **i64), "all_n": torch.empty(2 * world, **i64)}
    b = buf
    mark = mark or (lambda name: None)                                           # optional phase probe (tools/bench_sharded.py)
    backend.launch_transform(b["last"].data_ptr())
    dist.all_gather_into_tensor(b["all_dc"], b["last"], group=group)             # exchange 1: last DCs
    mark("transform")
    backend.launch_histogram(b["all_dc"].data_ptr() + 16 * (rank - 1) if rank else 0, b["hist"].data_ptr())
    dist.all_reduce(b["hist"], op=dist.ReduceOp.SUM, group=group)                # exchange 2: global histograms
    mark("histogram")
    backend.launch_tables(b["hist"].data_ptr(), b["bits"].data_ptr())
    dist.all_gather_into_tensor(b["all_bits"], b["bits"], group=group)           # exchange 3: bit counts
    mark("tables")
    torch.cumsum(b["all_bits"], 0, out=b["offs"])
    b["offs"].sub_(b["all_bits"])                                                # exclusive global bit offsets
    backend.launch_pack(b["offs"].data_ptr() + 8 * rank, rank == world - 1, b["tail"].data_ptr())
    dist.all_gather_into_tensor(b["all_tail"], b["tail"], group=group)           # trailing partial bytes
    mark("pack")
    d_bytes = backend.launch_stuff(b["all_tail"].data_ptr(), b["offs"].data_ptr(), b["all_bits"].data_ptr(), rank, world,
                                   b["n_bytes"].data_ptr())
    backend.launch_error(b["n_bytes"].data_ptr() + 8)                            # {byte count, device error flag}
    dist.all_gather_into_tensor(b["all_n"], b["n_bytes"], group=group)           # exchange 4: byte counts + status
    mark("stuff")
    res = b["all_n"].tolist()                                                    # the only host synchronisation
    mark("sync")
    _raise_first(res[1::2], "sharded encode (device-resident exchange)")         # the same error on every rank
    sizes = res[0::2]
    mine = (torch.as_tensor(_DevPtr(d_bytes, sizes[rank]), device=dev) if sizes[rank]
            else torch.empty(0, dtype=torch.uint8, device=dev))
    if rank == dst:
        out = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
        ops, off = [], 0
        for r in range(world):
            if sizes[r]:
                if r == rank:
                    out[off:off + sizes[r]].copy_(mine)
                else:
                    ops.append(dist.P2POp(dist.irecv, out[off:off + sizes[r]], r, group))
            off += sizes[r]
        if ops:
            for w in dist.batch_isend_irecv(ops):                                # all shards arrive concurrently
                w.wait()
        mark("gather")
        return out.cpu().numpy().tobytes() if to_host else out
    if mine.numel():
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, dst, group)]):
            w.wait()
    mark("gather")
    return None


class PeerFile:
    """The whole output file in the destination rank's HBM, mapped into every other rank's process through
    CUDA IPC (dmmt_peer_export / dmmt_peer_open), so that K4 of every shard stores its bytes at their final
    place over NVLink.  Collective: every rank of `group` constructs it once, after its CudaShardBackend."""

    def __init__(self, backend: CudaShardBackend, group=None, dst: int = 0, capacity: int | None = None):
        self.ctx, self.group, self.dst = backend.ctx, group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._ptr = C.c_void_p()
        self.capacity = 0
        self._allocate(int(capacity) if capacity else self._sum_of_strides(backend))

    def _sum_of_strides(self, backend: CudaShardBackend) -> int:
        strides = [None] * self.world
        dist.all_gather_object(strides, int(backend.out_stride), group=self.group)
        return sum(strides)

    def _allocate(self, capacity: int):
        self.capacity = capacity
        box = [None]
        if self.rank == self.dst:
            F.check(F.lib().dmmt_device_alloc(self.ctx.handle, self.capacity, C.byref(self._ptr)), "dmmt_device_alloc")
            h = C.create_string_buffer(64)
            F.check(F.lib().dmmt_peer_export(self.ctx.handle, self._ptr, h), "dmmt_peer_export")
            box[0] = h.raw
        src = dist.get_global_rank(self.group, self.dst) if self.group is not None else self.dst
        dist.broadcast_object_list(box, src=src, group=self.group)
        if self.rank != self.dst:
            F.check(F.lib().dmmt_peer_open(self.ctx.handle, box[0], C.byref(self._ptr)), "dmmt_peer_open")

    def regrow(self, backend: CudaShardBackend):
        """Collective: after the shards' scan capacity has grown, the file must hold the larger worst case too."""
        need = self._sum_of_strides(backend)
        if need > self.capacity:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)       # nobody still writes through the old mapping
            self.close()
            dist.barrier(group=self.group)       # every mapping is closed before the owner's next allocation is exported
            self._allocate(need)

    @property
    def ptr(self) -> int:
        return self._ptr.value

    def tensor(self, n: int) -> torch.Tensor:
        """first n bytes of the file as a uint8 tensor on the destination rank's device (no copy)"""
        if self.rank != self.dst:
            raise RuntimeError("the file lives on the destination rank")
        return torch.as_tensor(_DevPtr(self._ptr.value, n), device=f"cuda:{self.ctx.device}")

    def close(self):
        if self._ptr:
            if self.rank == self.dst:
                F.lib().dmmt_device_free(self.ctx.handle, self._ptr)
            else:
                F.lib().dmmt_peer_close(self.ctx.handle, self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerMailbox:
    """The exchanges of the sharded encode WITHOUT a collective library on the data path: every rank owns a mailbox in
    its HBM, mapped into the other processes through CUDA IPC (dmmt_mailbox_bytes, dmmt_peer_export / dmmt_peer_open);
    dmmt_shard_launch_post stores a rank's values into every mailbox over NVLink and releases them with the sequence
    number of the encode, dmmt_shard_launch_collect waits for all rows and sums / gathers them from local memory.
    torch.distributed is only used here, once, to hand the IPC handles round.  Collective: every rank of `group`
    constructs it after its CudaShardBackend and closes it with close()."""

    DC, HIST, BITS, TAIL, NBYTES, RES = range(6)          # slots: one per exchange of an encode
    SUM, GATHER, GATHER_SCAN = range(3)                   # modes of dmmt_shard_launch_collect

    def __init__(self, backend: CudaShardBackend, group=None):
        self.ctx, self.group = backend.ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.seq = 0
        self.bytes = int(F.lib().dmmt_mailbox_bytes(self.world))
        if not self.bytes:
            raise F.DmmtError(F.E_INVALID, "dmmt_mailbox_bytes")
        self._own = C.c_void_p()
        F.check(F.lib().dmmt_device_alloc(self.ctx.handle, self.bytes, C.byref(self._own)), "dmmt_device_alloc")
        torch.as_tensor(_DevPtr(self._own.value, self.bytes), device=f"cuda:{self.ctx.device}").zero_()
        torch.cuda.synchronize()
        h = C.create_string_buffer(64)
        F.check(F.lib().dmmt_peer_export(self.ctx.handle, self._own, h), "dmmt_peer_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, h.raw, group=group)
        self._ptrs = (C.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                self._ptrs[r] = self._own.value
            else:
                q = C.c_void_p()
                F.check(F.lib().dmmt_peer_open(self.ctx.handle, handles[r], C.byref(q)), "dmmt_peer_open")
                self._ptrs[r] = q.value
        dist.barrier(group=group)   # every mailbox is zeroed and mapped before anybody posts

    def begin(self) -> int:
        """sequence number of the next encode (the same on every rank: encodes are collective)"""
        self.seq += 1
        return self.seq

    def exchange(self, backend: CudaShardBackend, slot: int, src: torch.Tensor, n_words: int, mode: int, out: torch.Tensor):
        lib = F.lib()
        F.check(lib.dmmt_shard_launch_post(backend._h, self._ptrs, self.rank, self.world, slot, self.seq,
                                           C.c_void_p(src.data_ptr()), n_words), "dmmt_shard_launch_post")
        F.check(lib.dmmt_shard_launch_collect(backend._h, self._own, self.world, slot, self.seq, mode, n_words,
                                              C.c_void_p(out.data_ptr())), "dmmt_shard_launch_collect")

    def close(self):
        """Collective."""
        if not self._own:
            return
        torch.cuda.synchronize()
        for r in range(self.world):
            if r != self.rank and self._ptrs[r]:
                F.lib().dmmt_peer_close(self.ctx.handle, C.c_void_p(self._ptrs[r]))
                self._ptrs[r] = None
        dist.barrier(group=self.group)   # nobody maps this rank's mailbox any more
        F.lib().dmmt_device_free(self.ctx.handle, self._own)
        self._own = C.c_void_p()


def encode_sharded_peer(backend: CudaShardBackend, file: PeerFile, to_host: bool = True, mark=None,
                        mailbox: PeerMailbox | None = None):
    """see _encode_sharded_peer_once; DMMT_E_OVERFLOW on any shard: every rank grows its shard, runs once more"""
    return _retry_on_overflow(_encode_sharded_peer_once, backend, file, to_host, mark, mailbox)


def _encode_sharded_peer_once(backend: CudaShardBackend, file: PeerFile, to_host: bool = True, mark=None,
                              mailbox: PeerMailbox | None = None):
    """encode_sharded_device without the gather: after the tail exchange every shard counts its stuffed bytes
    (phase 5a), the counts are all-gathered and summed, and K4 (phase 5b) writes straight into `file` at the
    shard's final offset -- on the destination rank's own memory or, from the other ranks, over NVLink.  The
    closing all-gather of {end offset, error} is the completion barrier and the one host synchronisation.
    With a `mailbox` the six exchanges are the library's own kernels over peer memory (PeerMailbox); without
    one they are NCCL collectives on the same stream.
    Returns the file on the destination rank (bytes, or a view of `file` when to_host is False), None elsewhere."""
    group, dst = file.group, file.dst
    rank, world = file.rank, file.world
    dev = torch.device("cuda", backend.ctx.device)
    b = getattr(backend, "_pbuf", None)
    if b is None or b["world"] != world:
        i32, i64 = dict(dtype=torch.int32, device=dev), dict(dtype=torch.int64, device=dev)
        bits2, n2 = torch.empty(2 * world, **i64), torch.empty(2 * world, **i64)   # [values | exclusive prefix sums]
        b = backend._pbuf = {"world": world, "last": torch.empty(4, **i32), "all_dc": torch.empty(4 * world, **i32),
                             "hist": torch.empty(1024, **i64), "bits": torch.empty(1, **i64),
                             "bits2": bits2, "all_bits": bits2[:world], "offs": bits2[world:],
                             "tail": torch.empty(2, **i32), "all_tail": torch.empty(2 * world, **i32),
                             "n_bytes": torch.empty(1, **i64), "n2": n2, "all_n": n2[:world], "byte_offs": n2[world:],
                             "res": torch.empty(2, **i64), "all_res": torch.empty(2 * world, **i64)}
    mark = mark or (lambda name: None)
    mb = mailbox
    if mb is not None:
        mb.begin()
    backend.launch_transform(b["last"].data_ptr())
    if mb is not None:
        mb.exchange(backend, mb.DC, b["last"], 2, mb.GATHER, b["all_dc"])
    else:
        dist.all_gather_into_tensor(b["all_dc"], b["last"], group=group)         # exchange 1: last DCs
    mark("transform")
    backend.launch_histogram(b["all_dc"].data_ptr() + 16 * (rank - 1) if rank else 0, b["hist"].data_ptr())
    if mb is not None:
        mb.exchange(backend, mb.HIST, b["hist"], 1024, mb.SUM, b["hist"])
    else:
        dist.all_reduce(b["hist"], op=dist.ReduceOp.SUM, group=group)            # exchange 2: global histograms
    mark("histogram")
    backend.launch_tables(b["hist"].data_ptr(), b["bits"].data_ptr())
    if mb is not None:
        mb.exchange(backend, mb.BITS, b["bits"], 1, mb.GATHER_SCAN, b["bits2"])   # bit counts + exclusive bit offsets
    else:
        dist.all_gather_into_tensor(b["all_bits"], b["bits"], group=group)       # exchange 3: bit counts
        torch.cumsum(b["all_bits"], 0, out=b["offs"])
        b["offs"].sub_(b["all_bits"])                                            # exclusive global bit offsets
    mark("tables")
    backend.launch_pack(b["offs"].data_ptr() + 8 * rank, rank == world - 1, b["tail"].data_ptr())
    if mb is not None:
        mb.exchange(backend, mb.TAIL, b["tail"], 1, mb.GATHER, b["all_tail"])
    else:
        dist.all_gather_into_tensor(b["all_tail"], b["tail"], group=group)       # trailing partial bytes
    mark("pack")
    backend.launch_count_bytes(b["all_tail"].data_ptr(), b["offs"].data_ptr(), b["all_bits"].data_ptr(), rank, world,
                               b["n_bytes"].data_ptr())
    if mb is not None:
        mb.exchange(backend, mb.NBYTES, b["n_bytes"], 1, mb.GATHER_SCAN, b["n2"])  # byte counts + offsets in the file
    else:
        dist.all_gather_into_tensor(b["all_n"], b["n_bytes"], group=group)       # exchange 4: byte counts, BEFORE K4
        torch.cumsum(b["all_n"], 0, out=b["byte_offs"])
        b["byte_offs"].sub_(b["all_n"])                                          # exclusive byte offsets in the file
    mark("count")
    backend.launch_stuff_into(b["offs"].data_ptr(), rank, world, file.ptr, file.capacity,
                              b["byte_offs"].data_ptr() + 8 * rank, b["res"].data_ptr())
    if mb is not None:
        mb.exchange(backend, mb.RES, b["res"], 2, mb.GATHER, b["all_res"])
    else:
        dist.all_gather_into_tensor(b["all_res"], b["res"], group=group)         # completion barrier + status
    mark("stuff")
    res = b["all_res"].tolist()                                                  # the only host synchronisation
    mark("sync")
    _raise_first(res[1::2], "sharded encode (peer gather)")                      # the same error on every rank
    if rank != dst:
        return None
    out = file.tensor(int(res[2 * (world - 1)]))
    return out.cpu().numpy().tobytes() if to_host else out
