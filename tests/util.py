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


# ---- manifest / payload canonicalisation ---------------------------------------------------------
import hashlib
import json
import os
from typing import Any, Dict, Tuple


def _walk_tensor_entries(entry: Dict[str, Any]):
    t = entry.get("type")
    if t == "Tensor":
        yield entry
    elif t == "ChunkedTensor":
        for c in entry["chunks"]:
            yield c["tensor"]
    elif t in ("ShardedTensor", "DTensor"):
        for s in entry["shards"]:
            yield s["tensor"]


def canonicalize(manifest: Dict[str, Any]) -> Tuple[Dict[str, Any], Dict[str, str]]:
    """Slab files are named batched/<uuid4> (T:batcher.py:175): rename them batched/<k> by order of first
    appearance in manifest order.  Returns (canonical manifest, {original location -> canonical})."""
    manifest = json.loads(json.dumps(manifest))
    names: Dict[str, str] = {}
    for entry in manifest.values():
        for te in _walk_tensor_entries(entry):
            loc = te["location"]
            if loc.startswith("batched/"):
                if loc not in names:
                    names[loc] = f"batched/{len(names)}"
                te["location"] = names[loc]
    return manifest, names


def snapshot_digest(root: str) -> Dict[str, Any]:
    """{"manifest": canonical manifest, "files": {canonical location: {"nbytes", "sha256"}}} of a snapshot dir.
    torch_save payloads (pickles) are recorded as opaque: their bytes depend on the torch build."""
    meta = json.load(open(os.path.join(root, ".snapshot_metadata")))
    manifest, names = canonicalize(meta["manifest"])
    opaque = set()
    for entry in meta["manifest"].values():
        if entry.get("type") == "object":
            opaque.add(entry["location"])
        for te in _walk_tensor_entries(entry):
            if te["serializer"] != "buffer_protocol":
                opaque.add(te["location"])
    files: Dict[str, Any] = {}
    for dirpath, _, fnames in os.walk(root):
        for fn in fnames:
            full = os.path.join(dirpath, fn)
            rel = os.path.relpath(full, root)
            if rel == ".snapshot_metadata":
                continue
            canon = names.get(rel, rel)
            if rel in opaque:
                files[canon] = {"opaque": True}
            else:
                data = open(full, "rb").read()
                files[canon] = {"nbytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
    return {"version": meta["version"], "world_size": meta["world_size"], "manifest": manifest, "files": dict(sorted(files.items()))}
