"""A small, independent baseline-JPEG entropy DECODER (test infrastructure).

It follows ITU-T T.81 (Annex C code generation from the DHT counts, Annex F.2.2 decoding), not the
reference's encoder: marker segments -> quantisation tables, frame header, Huffman tables ->
de-stuffed scan -> DC/AC symbols -> quantised coefficients, zig-zag order, in MCU stream order.
Comparing its output with the oracle's quantised coefficient stream proves that the bytes our encoder
(and the oracle) write mean exactly those coefficients to any standard decoder.
"""
from __future__ import annotations

import numpy as np


class JpegError(ValueError):
    pass


def _huff_decode_table(counts, symbols):
    """T.81 Annex C: codes in order of increasing length, consecutive within a length."""
    table = {}
    code = 0
    k = 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


class _Bits:
    def __init__(self, data: bytes):
        self.data, self.pos, self.nbits, self.acc = data, 0, 0, 0

    def bit(self) -> int:
        if self.nbits == 0:
            if self.pos >= len(self.data):
                raise JpegError("scan exhausted")
            self.acc = self.data[self.pos]
            self.pos += 1
            self.nbits = 8
        self.nbits -= 1
        return (self.acc >> self.nbits) & 1

    def bits(self, n: int) -> int:
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v

    def symbol(self, table) -> int:
        code = 0
        for length in range(1, 17):
            code = (code << 1) | self.bit()
            s = table.get((length, code))
            if s is not None:
                return s
        raise JpegError("invalid Huffman code")

    def remaining_are_ones(self) -> bool:
        """the padding of the last byte must be 1-bits and nothing may follow"""
        ok = all(self.bit() for _ in range(self.nbits))
        return ok and self.pos == len(self.data)


def _extend(v: int, t: int) -> int:  # T.81 F.2.2.1
    return v if t == 0 or v >= (1 << (t - 1)) else v - (1 << t) + 1


def decode(jpeg: bytes):
    """-> dict(width, height, precision, sampling=[(h, v)...], qtables={id: zigzag 64}, dht={(cls, id): (counts, symbols)},
    coefficients=int16 [n_stream_blocks, 64] zig-zag, MCU-interleaved stream order, markers=[...])"""
    if jpeg[:2] != b"\xff\xd8":
        raise JpegError("no SOI")
    pos = 2
    q, dht, frame, scan_comps, markers = {}, {}, None, None, []
    while True:
        if jpeg[pos] != 0xFF:
            raise JpegError(f"marker expected at {pos}")
        m = jpeg[pos + 1]
        seglen = int.from_bytes(jpeg[pos + 2:pos + 4], "big")
        body = jpeg[pos + 4:pos + 2 + seglen]
        markers.append(m)
        pos += 2 + seglen
        if m == 0xDB:
            if body[0] >> 4:
                raise JpegError("16-bit DQT not expected")
            q[body[0] & 15] = list(body[1:65])
        elif m == 0xC0:
            frame = {"precision": body[0], "height": int.from_bytes(body[1:3], "big"),
                     "width": int.from_bytes(body[3:5], "big"),
                     "comps": [(body[6 + 3 * i], body[7 + 3 * i] >> 4, body[7 + 3 * i] & 15, body[8 + 3 * i])
                               for i in range(body[5])]}
        elif m == 0xC4:
            counts = list(body[1:17])
            dht[(body[0] >> 4, body[0] & 15)] = (counts, list(body[17:17 + sum(counts)]))
        elif m == 0xDA:
            n = body[0]
            scan_comps = [(body[1 + 2 * i], body[2 + 2 * i] >> 4, body[2 + 2 * i] & 15) for i in range(n)]
            if tuple(body[1 + 2 * n:4 + 2 * n]) != (0, 63, 0):
                raise JpegError("not a baseline sequential scan")
            break
    if jpeg[-2:] != b"\xff\xd9":
        raise JpegError("no EOI")
    raw = jpeg[pos:-2]
    # de-stuff: FF00 -> FF; any other FFxx inside the scan is an error for this encoder (no RSTn)
    out = bytearray()
    i = 0
    while i < len(raw):
        out.append(raw[i])
        if raw[i] == 0xFF:
            if i + 1 >= len(raw) or raw[i + 1] != 0x00:
                raise JpegError("unstuffed 0xFF in scan")
            i += 1
        i += 1
    bits = _Bits(bytes(out))
    hmax = max(c[1] for c in frame["comps"])
    vmax = max(c[2] for c in frame["comps"])
    mcus_x = -(-frame["width"] // (8 * hmax))
    mcus_y = -(-frame["height"] // (8 * vmax))
    tables = {k: _huff_decode_table(*v) for k, v in dht.items()}
    sel = {cid: (td, ta) for cid, td, ta in scan_comps}
    pred = {c[0]: 0 for c in frame["comps"]}
    blocks = []
    for _ in range(mcus_x * mcus_y):
        for cid, h, v, _tq in frame["comps"]:
            td, ta = sel[cid]
            for _b in range(h * v):
                zz = [0] * 64
                t = bits.symbol(tables[(0, td)])
                diff = _extend(bits.bits(t), t) if t else 0
                pred[cid] += diff
                zz[0] = pred[cid]
                k = 1
                while k < 64:
                    rs = bits.symbol(tables[(1, ta)])
                    r, s = rs >> 4, rs & 15
                    if s == 0:
                        if r == 15:
                            k += 16
                            continue
                        break  # EOB
                    k += r
                    if k > 63:
                        raise JpegError("AC run past the block")
                    zz[k] = _extend(bits.bits(s), s)
                    k += 1
                blocks.append(zz)
    if not bits.remaining_are_ones():
        raise JpegError("bad padding / trailing scan bytes")
    return {"width": frame["width"], "height": frame["height"], "precision": frame["precision"],
            "sampling": [(c[1], c[2]) for c in frame["comps"]], "tq": [c[3] for c in frame["comps"]],
            "qtables": q, "dht": dht, "markers": markers, "coefficients": np.array(blocks, dtype=np.int32)}
