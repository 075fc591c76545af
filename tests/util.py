"""Shared helpers for the test-suite: deterministic inputs and byte-level expectations."""
from __future__ import annotations

import torch

ALL_RAW_DTYPES = [
    torch.float64,
    torch.float32,
    torch.float16,
    torch.bfloat16,
    torch.int64,
    torch.int32,
    torch.int16,
    torch.int8,
    torch.uint8,
    torch.bool,
]


def det_bytes(n: int, seed: int) -> torch.Tensor:
    """n deterministic pseudo-random bytes (platform independent: integer arithmetic only)."""
    i = torch.arange(n, dtype=torch.int64)
    x = (i * 2654435761 + seed * 40503 + 12345) & 0xFFFFFFFF
    x = x ^ (x >> 13)
    x = (x * 1274126177) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    return (x & 0xFF).to(torch.uint8)


def det_tensor(shape, dtype: torch.dtype, seed: int) -> torch.Tensor:
    """Deterministic tensor of `dtype` with arbitrary bit patterns (bool: 0/1; floats: finite-ish raw bits)."""
    numel = 1
    for s in shape:
        numel *= s
    if dtype == torch.bool:
        return (det_bytes(numel, seed) & 1).to(torch.bool).reshape(shape)
    esz = torch.empty(0, dtype=dtype).element_size()
    raw = det_bytes(numel * esz, seed)
    return raw.view(dtype).reshape(shape).clone()


def wire_bytes(t: torch.Tensor) -> bytes:
    """The reference's payload for a view: C-contiguous native-endian element bytes."""
    t = t.detach().cpu().contiguous()
    if t.numel() == 0:
        return b""
    return bytes(t.reshape(-1).view(torch.uint8).numpy())


def same_bytes(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a.shape == b.shape and a.dtype == b.dtype and wire_bytes(a) == wire_bytes(b)
