"""Who writes what.  Replicated state is held by every rank; exactly one rank must write each piece.

Algorithm of the reference (T:partitioner.py:67-126, 140-213): every rank contributes
(entries, per-request sizes, its non-replicated byte count); rank 0 assigns each logical path — or, for
chunked tensors that are identical on all ranks, each chunk — to the currently least-loaded rank and
broadcasts the result.  Only ownership metadata crosses ranks; tensor bytes never move between GPUs."""
from __future__ import annotations

import copy
import hashlib
import pickle
import os
from collections import defaultdict
from dataclasses import dataclass
from typing import Dict, List, Set, Tuple

import numpy as np

from .io_preparers.object import ObjectBufferStager
from .io_preparers.tensor import TensorBufferStager, entry_nbytes
from .io_types import WriteReq
from .manifest import (
    ChunkedTensorEntry,
    DTensorEntry,
    Entry,
    is_fully_replicated_entry,
    is_partially_replicated_entry,
    is_replicated_entry,
)
from .pg_wrapper import PGWrapper


@dataclass(frozen=True)
class _WriteLoad:
    logical_path: str
    write_req_idx: int
    size: int


def replica_groups(entry: DTensorEntry) -> List[Set[int]]:
    """Sets of ranks that hold the same shard of a partially replicated DTensor: slices of the device
    mesh along its replicated mesh dims (T:manifest_utils.py:70-106)."""
    mesh = np.array(entry.mesh)
    sharded_mesh_dims = {m for dims in entry.dim_map if dims[0] != -1 for m in dims}
    groups: List[Set[int]] = []
    index_space = [range(n) if d in sharded_mesh_dims else [slice(None)] for d, n in enumerate(mesh.shape)]
    import itertools

    for idx in itertools.product(*index_space):
        groups.append({int(r) for r in np.asarray(mesh[idx]).reshape(-1)})
    return groups


def _write_size(wr: WriteReq) -> int:
    st = wr.buffer_stager
    if isinstance(st, TensorBufferStager):
        return entry_nbytes(st.entry)
    if isinstance(st, ObjectBufferStager):
        return st.get_staging_cost_bytes()
    raise AssertionError(f"Unrecognized buffer stager type {type(st)}")


def _partition_write_loads(
    rank_to_entries: List[Dict[str, Entry]],
    rank_to_write_loads: List[Dict[str, List[_WriteLoad]]],
    rank_to_size: List[int],
    world_size: int,
) -> List[List[_WriteLoad]]:
    result: List[List[_WriteLoad]] = [[] for _ in range(world_size)]
    load = list(rank_to_size)
    chunk_units: Set[_WriteLoad] = set()

    def give(candidates: List[int], path: str, size: int) -> None:
        r = min(candidates, key=lambda k: load[k])
        result[r].extend(rank_to_write_loads[r][path])
        load[r] += size

    for path, entry0 in rank_to_entries[0].items():
        same_everywhere = isinstance(entry0, ChunkedTensorEntry) and all(e[path] == entry0 for e in rank_to_entries)
        if same_everywhere:
            # chunk-granular: each chunk is its own unit of partitioning
            chunk_units.update(rank_to_write_loads[0][path])
            continue
        size = sum(wl.size for wl in rank_to_write_loads[0][path])
        if is_partially_replicated_entry(entry0):
            for group in replica_groups(entry0):  # type: ignore[arg-type]
                give(list(group), path, size)
        else:
            give(list(range(world_size)), path, size)
    # sorted: every rank derives the same assignment from the same gathered inputs (set order is hash-seed
    # dependent and differs between processes)
    for unit in sorted(chunk_units, key=lambda u: (u.logical_path, u.write_req_idx)):
        r = int(np.argmin(load))
        result[r].append(unit)
        load[r] += unit.size
    # callers pass rank_to_size by reference in the reference implementation; keep that contract
    rank_to_size[:] = load
    return result


def _partition_replicated_write_reqs(
    entries: Dict[str, Entry], write_reqs: Dict[str, List[WriteReq]], non_replicated_size: int, pg: PGWrapper
) -> Tuple[Dict[str, Entry], Dict[str, List[WriteReq]]]:
    loads: Dict[str, List[_WriteLoad]] = defaultdict(list)
    for path, wrs in write_reqs.items():
        for i, wr in enumerate(wrs):
            loads[path].append(_WriteLoad(path, i, _write_size(wr)))
    world = pg.get_world_size()
    # Replicated state is, by definition, the same on every rank (DDP): instead of shipping every rank's ~1000 entries to
    # every rank (O(world x entries) pickling per take: 18 ms of a 59 ms C2 take at 8 ranks), ranks first exchange a digest
    # of their replicated plan plus the one number that differs (their non-replicated bytes).  Only when the digests
    # disagree (partially replicated DTensors, diverging state) is the reference's full exchange needed
    # (T:partitioner.py:140-213).
    digest = hashlib.sha1(pickle.dumps((list(entries.items()), {k: [(w.write_req_idx, w.size) for w in v] for k, v in loads.items()}), protocol=4)).hexdigest()
    brief = [None] * world
    pg.all_gather_object(brief, (digest, non_replicated_size))
    if all(b[0] == digest for b in brief):
        all_entries, all_loads, all_sizes = [entries] * world, [loads] * world, [b[1] for b in brief]
    else:
        gathered = [None] * world
        pg.all_gather_object(gathered, (entries, loads, non_replicated_size))
        all_entries, all_loads, all_sizes = zip(*gathered)
    # The reference lets rank 0 partition and broadcasts the result (T:partitioner.py:176-192).  The greedy
    # assignment is a pure function of the gathered inputs, so every rank evaluates it locally: one collective less.
    result = _partition_write_loads(list(all_entries), list(all_loads), list(all_sizes), pg.get_world_size())
    mine = sorted((wl.logical_path, wl.write_req_idx) for wl in result[pg.get_rank()])
    new_entries: Dict[str, Entry] = {}
    new_reqs: Dict[str, List[WriteReq]] = defaultdict(list)
    for path, idx in mine:
        entry = entries[path]
        if isinstance(entry, ChunkedTensorEntry):
            if path not in new_entries:
                partial = copy.deepcopy(entry)
                partial.chunks = []
                new_entries[path] = partial
            new_entries[path].chunks.append(entry.chunks[idx])
        else:
            new_entries[path] = entry
        new_reqs[path].append(write_reqs[path][idx])
    return new_entries, new_reqs


def partition_write_reqs(
    entries: Dict[str, Entry], write_reqs: Dict[str, List[WriteReq]], pg: PGWrapper
) -> Tuple[Dict[str, Entry], Dict[str, List[WriteReq]]]:
    missing = set(write_reqs) - set(entries)
    if missing:
        raise RuntimeError(f"Not all entries associated with the write reqs are passed in. Missing: {missing}.")
    if os.environ.get("TORCH_SNAPSHOT_DISABLE_PARTITIONER") is not None:
        raise NotImplementedError("TORCH_SNAPSHOT_DISABLE_PARTITIONER is not implemented.")
    rep_entries = {k: v for k, v in entries.items() if is_replicated_entry(v)}
    rep_reqs = {k: v for k, v in write_reqs.items() if k in rep_entries}
    own_entries = {k: v for k, v in entries.items() if k not in rep_entries}
    own_reqs = {k: v for k, v in write_reqs.items() if k in own_entries}
    own_bytes = sum(_write_size(wr) for wrs in own_reqs.values() for wr in wrs)
    rep_entries, rep_reqs = _partition_replicated_write_reqs(rep_entries, rep_reqs, own_bytes, pg)
    return {**rep_entries, **own_entries}, {**rep_reqs, **own_reqs}


def consolidate_replicated_entries(rank_to_entries: List[Dict[str, Entry]], dedup: bool = True) -> List[Dict[str, Entry]]:
    """After partitioning every rank only knows the pieces it writes; stitch chunked entries back
    together and keep fully replicated entries in rank 0's manifest only (T:partitioner.py:285-355)."""
    by_path: Dict[str, List[ChunkedTensorEntry]] = defaultdict(list)
    for entries in rank_to_entries:
        for path, e in entries.items():
            if isinstance(e, ChunkedTensorEntry) and is_replicated_entry(e):
                by_path[path].append(e)
    for path, parts in by_path.items():
        merged = ChunkedTensorEntry(
            dtype=parts[0].dtype,
            shape=parts[0].shape,
            chunks=sorted((c for p in parts for c in p.chunks), key=lambda c: c.offsets),
            replicated=True,
        )
        for entries in rank_to_entries:
            entries[path] = merged
    shared: Dict[str, Entry] = {}
    for entries in rank_to_entries:
        for path in list(entries):
            e = entries[path]
            if not is_fully_replicated_entry(e):
                continue
            if path in shared and shared[path] != e:
                raise ValueError(f"Paths for replicated entry for {path} do not match: {shared[path]} vs. {e}")
            shared.setdefault(path, e)
            del entries[path]
    for rank, entries in enumerate(rank_to_entries):
        if dedup and rank != 0:
            continue
        entries.update(shared)
    return rank_to_entries


def consolidate_replicated_entries_dist(entries: Dict[str, Entry], pg: PGWrapper, dedup: bool = True) -> Dict[str, Entry]:
    """Collective form of :func:`consolidate_replicated_entries` (T:partitioner.py:358-368)."""
    gathered: List[Dict[str, Entry]] = [None] * pg.get_world_size()  # type: ignore[list-item]
    pg.all_gather_object(gathered, entries)
    return consolidate_replicated_entries(gathered, dedup=dedup)[pg.get_rank()]
