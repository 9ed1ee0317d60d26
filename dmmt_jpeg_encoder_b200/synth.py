"""Deterministic synthetic workloads for bench.py (measurement infrastructure, not the product).

`photo(index, h, w, device)` is a photo-like frame (SURVEY 8d "photo": low-pass noise mixed with a
diagonal ramp plus fine noise) built ONLY from int64 torch ops, so the CPU and the CUDA device
produce bit-identical pixels: the GPU arm and the CPU-baseline arm time the same images.
`grad` is the reference's own dct_timing pattern (bin/dct_timing.rs:150-160) extended to 3 channels.
"""
from __future__ import annotations

import torch

_M32 = 0xFFFFFFFF


def _hash32(s: torch.Tensor) -> torch.Tensor:
    z = s & _M32
    z = ((z ^ (z >> 15)) * 0x2C1B3C6D) & _M32
    z = ((z ^ (z >> 12)) * 0x297A2D39) & _M32
    return z ^ (z >> 15)


def photo(index: int, h: int, w: int, device="cpu") -> torch.Tensor:
    """-> uint8 [h, w, 3]."""
    dev = torch.device(device)
    y = torch.arange(h, dtype=torch.int64, device=dev).view(h, 1, 1)
    x = torch.arange(w, dtype=torch.int64, device=dev).view(1, w, 1)
    c = torch.arange(3, dtype=torch.int64, device=dev).view(1, 1, 3)
    flat = (y * w + x) * 3 + c
    salt = (1234 + index) * 2654435761
    n = _hash32(flat + salt) & 0xFFFF                       # uniform 16-bit noise
    for _ in range(3):                                      # 3 x 5-point blur (wrap-around edges)
        n = (n + torch.roll(n, 1, 0) + torch.roll(n, -1, 0) + torch.roll(n, 1, 1) + torch.roll(n, -1, 1)) // 5
    # 16.16 fixed point: 0.6 * (127.5 + 63.75 * (n - 32770) / 4787) + 0.4 * 255 * (x + y) / (w + h) + fine noise
    low = (n - 32770) * ((6375 << 16) // (100 * 4787))      # 63.75 levels per sigma (4787), scaled by 2^16
    base = (low * 6) // 10 + ((765 << 16) // 10)            # 0.6 * 127.5 = 76.5
    ramp = ((x + y) * (102 << 16)) // (w + h)               # 0.4 * 255 = 102
    fine = ((_hash32(flat * 7 + salt + 0x5BD1E995) & 0xFF) - 128) * ((3 << 16) // 128)  # about +-3 levels
    v = (base + ramp + fine + (1 << 15)) >> 16
    return v.clamp_(0, 255).to(torch.uint8)


def grad(h: int, w: int, device="cpu", y0: int = 0) -> torch.Tensor:
    """rows [y0, y0 + h) of the pattern (position-pure, so MCU-row shards can generate their own rows)"""
    dev = torch.device(device)
    y = torch.arange(y0, y0 + h, dtype=torch.int64, device=dev).view(h, 1)
    x = torch.arange(w, dtype=torch.int64, device=dev).view(1, w)
    return torch.stack([(x + 8 * y) & 255, (2 * x + 3 * y + 85) & 255, (5 * x + y + 170) & 255], -1).to(torch.uint8)


def uniform(index: int, h: int, w: int, device="cpu", y0: int = 0) -> torch.Tensor:
    dev = torch.device(device)
    flat = torch.arange(y0 * w * 3, (y0 + h) * w * 3, dtype=torch.int64, device=dev).view(h, w, 3)
    return (_hash32(flat + (1234 + index) * 2654435761) & 0xFF).to(torch.uint8)


def smooth(index: int, h: int, w: int, device="cpu", y0: int = 0) -> torch.Tensor:
    """Position-pure photo-like rows [y0, y0 + h): three incommensurate integer ramps per channel plus
    +-32 levels of hash noise (about 0.11 B/px at 4:2:0); used for the MCU-row sharded workload."""
    dev = torch.device(device)
    y = torch.arange(y0, y0 + h, dtype=torch.int64, device=dev).view(h, 1, 1)
    x = torch.arange(w, dtype=torch.int64, device=dev).view(1, w, 1)
    c = torch.arange(3, dtype=torch.int64, device=dev).view(1, 1, 3)
    tri = lambda t, p: ((t % (2 * p)) - p).abs()                     # triangle wave 0..p
    v = (tri(x * (3 + c) + y * 2, 1531) * 255) // 1531 + (tri(y * (5 + 2 * c) + x, 977) * 255) // 977
    v = v // 2 + ((_hash32((y * w + x) * 3 + c + (1234 + index) * 2654435761) & 63) - 32)
    return v.clamp_(0, 255).to(torch.uint8)


def make(kind: str, index: int, h: int, w: int, device="cpu", y0: int = 0) -> torch.Tensor:
    if kind == "photo":
        assert y0 == 0, "photo is not position-pure (blur wraps around the whole frame)"
        return photo(index, h, w, device)
    if kind == "grad":
        return grad(h, w, device, y0)
    if kind == "uniform":
        return uniform(index, h, w, device, y0)
    if kind == "smooth":
        return smooth(index, h, w, device, y0)
    raise ValueError(kind)
