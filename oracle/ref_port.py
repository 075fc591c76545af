"""ORACLE — test infrastructure, not product code.

A CPU restatement (numpy for the byte/index arithmetic) of pytorch/torchsnapshot's save/restore hot
path, used ONLY as a checker by ``tests/``, by ``__graft_entry__.smoke()`` and as the timed
``cpu_baseline`` / ``--impl reference`` arm of ``bench.py``.  Nothing under ``torchsnapshot_b200/``
imports it.

Pinned: ``tests/test_oracle.py`` checks this module against (a) the golden manifests + payload
digests produced by the unmodified reference in this container (``oracle/gen_golden.py`` ->
``tests/golden/*.json``), (b) the known-answer chunk/shard plans asserted by the reference's own
tests (tests/test_chunked_tensor_io_preparer.py:52-103, tests/test_sharded_tensor_io_preparer.py:212-297),
and (c) the live reference when ``/root/reference`` is present.

Each function cites the reference lines it restates ("T:" = torchsnapshot/ in the reference tree).
"""
from __future__ import annotations

import asyncio
import json
import math
import os
import threading
import uuid
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

MiB = 1024 * 1024
DEFAULT_MAX_CHUNK = 512 * MiB  # T:knobs.py:29
DEFAULT_MAX_SHARD = 512 * MiB  # T:knobs.py:30
DEFAULT_SLAB_THRESHOLD = 128 * MiB  # T:knobs.py:31

# T:serialization.py:162-173
RAW_DTYPES = {
    "torch.float64": 8, "torch.float32": 4, "torch.float16": 2, "torch.bfloat16": 2,
    "torch.int64": 8, "torch.int32": 4, "torch.int16": 2, "torch.int8": 1, "torch.uint8": 1, "torch.bool": 1,
}  # fmt: skip


# --------------------------------------------------------------------------------------------------
# index arithmetic
# --------------------------------------------------------------------------------------------------
def chunk_plan(shape: Sequence[int], itemsize: int, max_chunk_bytes: int = DEFAULT_MAX_CHUNK, dim: int = 0):
    """[(offsets, sizes)] of the dim-0 chunks of a tensor.  T:io_preparers/chunked_tensor.py:36-64:
    0-d -> 1-d; n = ceil(bytes / limit); torch.chunk(t, n, dim) => every chunk has ceil(extent / n) rows
    except a ragged last one (and fewer than n chunks may come out)."""
    shape = [1] if len(shape) == 0 else list(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * itemsize
    n = math.ceil(nbytes / max_chunk_bytes)
    extent = shape[dim]
    rows = math.ceil(extent / n)
    out = []
    for lo in range(0, extent, rows):
        off = [0] * len(shape)
        off[dim] = lo
        sz = list(shape)
        sz[dim] = min(rows, extent - lo)
        out.append((off, sz))
    return out


def subdivide_plan(offsets: Sequence[int], sizes: Sequence[int], dim: int, itemsize: int, max_shard_bytes: int = DEFAULT_MAX_SHARD):
    """[(local_start, offsets, sizes)] of the pieces of one local shard.  T:io_preparers/sharded_tensor.py:48-78:
    slice_sz = prod(sizes) / sizes[dim] * itemsize; chunk_length = max(floor(limit / slice_sz), 1)."""
    if max_shard_bytes <= 0:
        raise ValueError("max_shard_sz_bytes must be a positive integer")
    slice_sz = int(np.prod(sizes, dtype=np.int64)) // sizes[dim] * itemsize
    length = max(math.floor(max_shard_bytes / slice_sz), 1)
    n = math.ceil(sizes[dim] / length)
    out = []
    for i in range(n):
        start = i * length
        cur = min((i + 1) * length, sizes[dim]) - start
        off = list(offsets)
        off[dim] += start
        sz = list(sizes)
        sz[dim] = cur
        out.append((start, off, sz))
    return out


def boxes_overlap(a_off, a_sz, b_off, b_sz) -> bool:
    """torch.distributed._shard.sharding_spec._internals._check_shard_metadata_pair_overlap, as called
    at T:io_preparers/sharded_tensor.py:239: overlap unless separated along some dim."""
    for ao, asz, bo, bsz in zip(a_off, a_sz, b_off, b_sz):
        if ao >= bo + bsz or bo >= ao + asz:
            return False
    return True


def overlap_region(saved_off, saved_sz, cur_off, cur_sz):
    """[(dim, offset_in_saved, offset_in_current, length)].  T:io_preparers/sharded_tensor.py:80-127."""
    out = []
    for d, (so, co, ss, cs) in enumerate(zip(saved_off, cur_off, saved_sz, cur_sz)):
        end = min(so + ss, co + cs)
        length = end - max(co, so)
        if so > co:
            s_rel, c_rel = 0, so - co
        else:
            s_rel, c_rel = co - so, 0
        out.append((d, s_rel, c_rel, length))
    return out


def slab_assign(reqs: Sequence[Tuple[str, int, bool, bool]], threshold: int = DEFAULT_SLAB_THRESHOLD):
    """First-fit slab assignment.  T:batcher.py:246-319.

    reqs: (path, nbytes, is_cuda, batchable) in write-request order.
    Returns (passthrough_paths, slabs) with slabs = [{"cuda": bool, "members": [(path, lo, hi)]}] in the
    reference's emission order (all CPU slabs, then all GPU slabs; empty slabs dropped)."""
    passthrough: List[str] = []
    chains = {False: [[]], True: [[]]}
    sizes = {False: [0], True: [0]}
    for path, nbytes, is_cuda, batchable in reqs:
        if not batchable or nbytes >= threshold:
            passthrough.append(path)
            continue
        chain, sz = chains[is_cuda], sizes[is_cuda]
        if sz[-1] + nbytes >= threshold:
            chain.append([])
            sz.append(0)
        chain[-1].append((path, sz[-1], sz[-1] + nbytes))
        sz[-1] += nbytes
    slabs = [{"cuda": c, "members": m} for c in (False, True) for m in chains[c] if m]
    return passthrough, slabs


def partition_greedy(paths_sizes: Sequence[Tuple[str, int]], rank_loads: Sequence[int]) -> List[List[str]]:
    """Non-subpartitionable replicated paths go, in order, to the least-loaded rank (first minimum).
    T:partitioner.py:51-64 (`min(ranks, key=load)`), 108-113."""
    loads = list(rank_loads)
    out: List[List[str]] = [[] for _ in loads]
    for path, size in paths_sizes:
        r = min(range(len(loads)), key=lambda k: loads[k])
        out[r].append(path)
        loads[r] += size
    return out


# --------------------------------------------------------------------------------------------------
# byte arithmetic
# --------------------------------------------------------------------------------------------------
def _storage_bytes(t: torch.Tensor) -> np.ndarray:
    st = t.untyped_storage()
    if st.nbytes() == 0:
        return np.zeros(0, dtype=np.uint8)
    return torch.empty(0, dtype=torch.uint8).set_(st).numpy()


def serialize_view(t: torch.Tensor) -> bytes:
    """Payload of one tensor view under the buffer-protocol serializer: C-order bytes of the logical
    elements, native endianness, no header.  T:serialization.py:177-204 (`.contiguous()` then
    `memoryview(...).cast("b")`), 208-230 for bfloat16.  Restated with numpy on the raw storage so that it
    is dtype-agnostic (void elements) and independent of torch's copy kernels."""
    t = t.detach()
    if t.device.type != "cpu":
        t = t.cpu()  # checker-side only
    if t.numel() == 0:
        return b""
    isz = t.element_size()
    raw = _storage_bytes(t)[t.storage_offset() * isz :]
    # elements as (…, isz) byte rows: works for every dtype, bfloat16 included
    view = np.lib.stride_tricks.as_strided(
        raw, shape=tuple(t.shape) + (isz,), strides=tuple(st * isz for st in t.stride()) + (1,), writeable=False
    )
    return np.ascontiguousarray(view).tobytes()


def box(t: torch.Tensor, offsets: Sequence[int], sizes: Sequence[int]) -> torch.Tensor:
    """T:io_preparers/chunked_tensor.py:66-75 (_get_subtensor_view)."""
    v = t.view(-1) if t.dim() == 0 else t
    for d in range(len(sizes)):
        v = v.narrow(d, offsets[d], sizes[d])
    return v


def scatter_bytes(buf: bytes, saved_shape: Sequence[int], dtype: torch.dtype, regions, dst: torch.Tensor) -> None:
    """Restore-side restatement: frombuffer -> reshape -> narrow both sides -> copy_.
    T:io_preparers/tensor.py:319-340, T:io_preparers/sharded_tensor.py:285-323.  `regions` is a list of
    overlap_region() results against views of `dst` (CPU tensor)."""
    isz = torch.empty(0, dtype=dtype).element_size()
    src = np.frombuffer(buf, dtype=np.uint8).reshape(tuple(saved_shape) + (isz,))
    raw = _storage_bytes(dst)[dst.storage_offset() * isz :]
    full = np.lib.stride_tricks.as_strided(raw, shape=tuple(dst.shape) + (isz,), strides=tuple(st * isz for st in dst.stride()) + (1,))
    for region in regions:
        s_idx = tuple(slice(so, so + n) for _, so, _, n in region)
        d_idx = tuple(slice(do, do + n) for _, _, do, n in region)
        full[d_idx] = src[s_idx]


# --------------------------------------------------------------------------------------------------
# plan of one rank's save (world_size == 1 or per-rank non-replicated state)
# --------------------------------------------------------------------------------------------------
def dtype_str(dt: torch.dtype) -> str:
    return str(dt)


def tensor_entry(location: str, t: torch.Tensor, replicated: bool) -> Dict[str, Any]:
    """T:io_preparers/tensor.py:50-89, field order of T:manifest.py:49-90."""
    raw = dtype_str(t.dtype) in RAW_DTYPES
    return {
        "type": "Tensor",
        "location": location,
        "serializer": "buffer_protocol" if raw else "torch_save",
        "dtype": dtype_str(t.dtype),
        "shape": list(t.shape),
        "replicated": replicated,
        "byte_range": None,
    }


class ShardedSpec:
    """Description of a ShardedTensor's local shards for the oracle: [(tensor, offsets, sizes)], sharding dim."""

    def __init__(self, local_shards: List[Tuple[torch.Tensor, List[int], List[int]]], dim: int = 0) -> None:
        self.local_shards = local_shards
        self.dim = dim


def plan_save(
    flattened: Dict[str, Any],
    rank: int = 0,
    replicated_paths: Iterable[str] = (),
    max_chunk: int = DEFAULT_MAX_CHUNK,
    max_shard: int = DEFAULT_MAX_SHARD,
    slab_threshold: int = DEFAULT_SLAB_THRESHOLD,
    batching: bool = True,
) -> Tuple[Dict[str, Dict[str, Any]], Dict[str, bytes]]:
    """Entries (json-able, keyed by logical path) and payload files {location: bytes} that the reference
    produces for the raw-tensor leaves of `flattened` on one rank.  Restates T:snapshot.py:584-616 for
    tensors / ShardedSpec leaves: prepare_write (T:io_preparer.py:82-147, storage paths :52-61) ->
    batch_write_requests (T:batcher.py:204-355).  Slab files are named ``batched/<k>`` in emission order
    (the reference uses uuid4; tests canonicalise).  Non-raw leaves are ignored."""
    replicated_paths = set(replicated_paths)
    entries: Dict[str, Dict[str, Any]] = {}
    reqs: List[Tuple[str, torch.Tensor, Dict[str, Any]]] = []  # (location, view, tensor entry)
    for lp, obj in flattened.items():
        rep = lp in replicated_paths
        if isinstance(obj, ShardedSpec):
            loc = os.path.join("replicated_sharded" if rep else "sharded", lp)
            shards = []
            for t, off, sz in obj.local_shards:
                for start, p_off, p_sz in subdivide_plan(off, sz, obj.dim, t.element_size(), max_shard):
                    view = t.narrow(obj.dim, start, p_sz[obj.dim])
                    e = tensor_entry(f"{loc}_{'_'.join(str(i) for i in p_off)}", view, False)
                    shards.append({"offsets": p_off, "sizes": p_sz, "tensor": e})
                    reqs.append((e["location"], view, e))
            entries[lp] = {"type": "ShardedTensor", "shards": shards}
        elif isinstance(obj, torch.Tensor):
            loc = os.path.join("replicated" if rep else str(rank), lp)
            if obj.numel() * obj.element_size() > max_chunk:
                chunks = []
                for off, sz in chunk_plan(list(obj.shape), obj.element_size(), max_chunk):
                    view = box(obj, off, sz)
                    e = tensor_entry(f"{loc}_{'_'.join(str(i) for i in off)}", view, False)
                    chunks.append({"offsets": off, "sizes": sz, "tensor": e})
                    reqs.append((e["location"], view, e))
                entries[lp] = {"type": "ChunkedTensor", "dtype": dtype_str(obj.dtype), "shape": list(obj.shape), "chunks": chunks, "replicated": rep}
            else:
                e = tensor_entry(loc, obj, rep)
                entries[lp] = e
                reqs.append((loc, obj, e))
    files: Dict[str, bytes] = {}
    raw_reqs = [(loc, v, e) for loc, v, e in reqs if e["serializer"] == "buffer_protocol"]
    if batching:
        order = [(loc, v.numel() * v.element_size(), bool(v.is_cuda), True) for loc, v, _ in raw_reqs]
        passthrough, slabs = slab_assign(order, slab_threshold)
        by_loc = {loc: (v, e) for loc, v, e in raw_reqs}
        for loc in passthrough:
            files[loc] = serialize_view(by_loc[loc][0])
        for k, slab in enumerate(slabs):
            name = f"batched/{k}"
            blob = bytearray()
            for loc, lo, hi in slab["members"]:
                v, e = by_loc[loc]
                payload = serialize_view(v)
                assert len(payload) == hi - lo and len(blob) == lo
                blob += payload
                e["location"] = name
                e["byte_range"] = [lo, hi]
            files[name] = bytes(blob)
    else:
        for loc, v, _ in raw_reqs:
            files[loc] = serialize_view(v)
    return entries, files


# --------------------------------------------------------------------------------------------------
# the reference's execution pipeline, restated for timing (cpu_baseline / --impl reference)
# --------------------------------------------------------------------------------------------------
class RefPipeline:
    """Restates how the reference *executes* a save/restore of raw tensors so that it can be timed on the
    same box: T:scheduler.py:222-339 (asyncio; stage in a 4-thread pool, <=16 concurrent writes, return
    when staged, drain in complete()), T:io_preparers/tensor.py:240-271 + :353-355 (`tensor.to("cpu")`,
    pageable), T:batcher.py:144-159 (GPU slab: uint8 tensor, per-member `.contiguous()` + copy into
    [lo,hi), then `.cpu()` on the loop thread), T:batcher.py:66-93 (CPU slab: bytearray + slice assign),
    T:storage_plugins/fs.py:28-38 (one thread hop per file, open "wb+", single write, no fsync), and on
    restore T:storage_plugins/fs.py:40-51 + T:scheduler.py:369-376 (read -> BytesIO -> getvalue) +
    T:io_preparers/tensor.py:331-340 (frombuffer + copy_ in the 4-thread pool)."""

    CPU_THREADS = 4  # T:scheduler.py:32
    IO_CONCURRENCY = 16  # T:knobs.py:38

    def __init__(self, root: str, slab_threshold: int = DEFAULT_SLAB_THRESHOLD, max_chunk: int = DEFAULT_MAX_CHUNK) -> None:
        self.root = root
        self.slab_threshold = slab_threshold
        self.max_chunk = max_chunk

    # -- planning (same rules as plan_save, but keeps tensors instead of bytes) --
    def _plan(self, tensors: Dict[str, torch.Tensor]):
        reqs: List[Tuple[str, torch.Tensor]] = []
        layout: Dict[str, List[Tuple[str, List[int], List[int]]]] = {}
        for lp, t in tensors.items():
            loc = os.path.join("0", lp)
            if t.numel() * t.element_size() > self.max_chunk:
                layout[lp] = []
                for off, sz in chunk_plan(list(t.shape), t.element_size(), self.max_chunk):
                    name = f"{loc}_{'_'.join(str(i) for i in off)}"
                    reqs.append((name, box(t, off, sz)))
                    layout[lp].append((name, off, sz))
            else:
                reqs.append((loc, t))
                layout[lp] = [(loc, [0] * t.dim(), list(t.shape))]
        order = [(loc, v.numel() * v.element_size(), bool(v.is_cuda), True) for loc, v in reqs]
        passthrough, slabs = slab_assign(order, self.slab_threshold)
        return dict(reqs), passthrough, slabs, layout

    def save(self, tensors: Dict[str, torch.Tensor]) -> Dict[str, Any]:
        by_loc, passthrough, slabs, layout = self._plan(tensors)
        index: Dict[str, Any] = {"layout": layout, "where": {}}
        for loc in passthrough:
            index["where"][loc] = (loc, None)
        for k, slab in enumerate(slabs):
            for loc, lo, hi in slab["members"]:
                index["where"][loc] = (f"batched/{k}", (lo, hi))
        loop = asyncio.new_event_loop()
        try:
            loop.run_until_complete(self._save_async(by_loc, passthrough, slabs))
        finally:
            loop.close()
        return index

    async def _save_async(self, by_loc, passthrough, slabs) -> None:
        loop = asyncio.get_running_loop()
        stage_pool = ThreadPoolExecutor(max_workers=self.CPU_THREADS)
        io_pool = ThreadPoolExecutor(max_workers=self.IO_CONCURRENCY)
        io_gate = asyncio.Semaphore(self.IO_CONCURRENCY)

        def to_cpu(t: torch.Tensor) -> memoryview:
            c = t.detach().to("cpu") if t.is_cuda else t.detach()
            c = c.contiguous()
            return memoryview(c.reshape(-1).view(torch.uint8).numpy()) if c.numel() else memoryview(b"")

        def write_file(path: str, buf) -> None:
            full = os.path.join(self.root, path)
            os.makedirs(os.path.dirname(full), exist_ok=True)
            with open(full, "wb+") as f:
                f.write(buf)

        async def write(path: str, buf) -> None:
            async with io_gate:
                await loop.run_in_executor(io_pool, write_file, path, buf)

        async def stage(t: torch.Tensor):
            # CUDA tensors hop to the 4-thread pool for the D2H copy; CPU tensors are viewed in place on the loop
            # thread (T:io_preparers/tensor.py:249-266)
            if t.is_cuda:
                return await loop.run_in_executor(stage_pool, to_cpu, t)
            return to_cpu(t)

        async def single(loc: str) -> None:
            buf = await stage(by_loc[loc])
            await write(loc, buf)

        async def slab_task(k: int, slab) -> None:
            size = slab["members"][-1][2]
            if slab["cuda"]:
                gpu = torch.empty(size, dtype=torch.uint8, device=by_loc[slab["members"][0][0]].device)
                for loc, lo, hi in slab["members"]:
                    src = by_loc[loc].detach().contiguous()
                    if src.numel():
                        gpu[lo:hi].copy_(src.reshape(-1).view(torch.uint8))
                host = gpu.cpu()  # on the event-loop thread, as the reference does
                buf = memoryview(host.numpy())
            else:
                blob = bytearray(size)
                staged = await asyncio.gather(*[stage(by_loc[loc]) for loc, _, _ in slab["members"]])
                for (loc, lo, hi), b in zip(slab["members"], staged):
                    blob[lo:hi] = b
                buf = memoryview(blob)
            await write(f"batched/{k}", buf)

        tasks = [single(loc) for loc in passthrough] + [slab_task(k, s) for k, s in enumerate(slabs)]
        await asyncio.gather(*tasks)
        stage_pool.shutdown()
        io_pool.shutdown()

    def load(self, index: Dict[str, Any], out: Dict[str, torch.Tensor]) -> None:
        loop = asyncio.new_event_loop()
        try:
            loop.run_until_complete(self._load_async(index, out))
        finally:
            loop.close()

    async def _load_async(self, index, out) -> None:
        loop = asyncio.get_running_loop()
        cpu_pool = ThreadPoolExecutor(max_workers=self.CPU_THREADS)
        io_pool = ThreadPoolExecutor(max_workers=self.IO_CONCURRENCY)
        io_gate = asyncio.Semaphore(self.IO_CONCURRENCY)
        # merge ranged reads per file (T:batcher.py:387-478)
        per_file: Dict[str, List[Tuple[Optional[Tuple[int, int]], torch.Tensor]]] = {}
        for lp, pieces in index["layout"].items():
            for name, off, sz in pieces:
                path, br = index["where"][name]
                per_file.setdefault(path, []).append((br, box(out[lp], off, sz)))

        def read_file(path: str, br):
            import io as _io

            with open(os.path.join(self.root, path), "rb") as f:
                if br is None:
                    data = f.read()
                else:
                    f.seek(br[0])
                    data = f.read(br[1] - br[0])
            return _io.BytesIO(data).getvalue()

        def consume(dst: torch.Tensor, buf) -> None:
            if dst.numel():
                src = torch.frombuffer(buf, dtype=dst.dtype).reshape(dst.shape)
                dst.detach().copy_(src)

        async def one(path: str, members) -> None:
            if members[0][0] is None:
                async with io_gate:
                    buf = await loop.run_in_executor(io_pool, read_file, path, None)
                await loop.run_in_executor(cpu_pool, consume, members[0][1], buf)
                return
            lo = min(br[0] for br, _ in members)
            hi = max(br[1] for br, _ in members)
            async with io_gate:
                buf = await loop.run_in_executor(io_pool, read_file, path, (lo, hi))
            await asyncio.gather(*[loop.run_in_executor(cpu_pool, consume, dst, bytearray(buf[br[0] - lo : br[1] - lo])) for br, dst in members])

        await asyncio.gather(*[one(p, m) for p, m in per_file.items()])
        cpu_pool.shutdown()
        io_pool.shutdown()
