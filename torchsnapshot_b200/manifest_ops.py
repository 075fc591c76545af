"""Per-rank views of the global manifest and restore-time elasticity (T:manifest_ops.py:35-287).

``get_manifest_for_rank``: replicated entries (stored under rank 0 only) become visible to every rank;
ShardedTensor / DTensor entries are replaced by the union of all ranks' shards, so any rank can load
any region (reshard-on-load).  Ranks beyond the saved world size see only replicated state."""
from __future__ import annotations

import copy
from collections import defaultdict
from typing import Dict, List, Set, Tuple

from .knobs import is_sharded_tensor_elasticity_enabled_at_root_only
from .manifest import (
    DTensorEntry,
    Entry,
    Manifest,
    ShardedTensorEntry,
    SnapshotMetadata,
    is_container_entry,
    is_dict_entry,
    is_fully_replicated_entry,
)
from .partitioner import replica_groups


def _split_by_rank(metadata: SnapshotMetadata) -> List[Dict[str, Entry]]:
    """Per-rank views.  Entries are shared with ``metadata`` (read-only here); only container entries, whose
    key lists the elasticity pass edits in place, are copied — a deepcopy of the whole manifest costs
    O(world x entries) Python work per restore (≈0.3 s for Llama-3-8B saved at 8 ranks)."""
    per_rank: List[Dict[str, Entry]] = [{} for _ in range(metadata.world_size)]
    for path, entry in metadata.manifest.items():
        rank, _, logical = path.partition("/")
        per_rank[int(rank)][logical] = copy.deepcopy(entry) if is_container_entry(entry) else entry
    return per_rank


def _merge_sharded(per_rank: List[Dict[str, Entry]]) -> Dict[str, Entry]:
    shards = defaultdict(list)
    for m in per_rank:
        for path, e in m.items():
            if isinstance(e, ShardedTensorEntry):
                shards[path].extend(e.shards)
    return {p: ShardedTensorEntry(shards=sorted(s, key=lambda sh: sh.offsets)) for p, s in shards.items()}


def _merge_dtensors(per_rank: List[Dict[str, Entry]]) -> Dict[str, Entry]:
    """Union of shards over ranks, taking each replicated shard from one rank of its replica group only."""
    parts: Dict[str, List[DTensorEntry]] = defaultdict(list)
    covered: Dict[str, Set[int]] = defaultdict(set)
    groups: Dict[str, List[Set[int]]] = {}
    for rank, m in enumerate(per_rank):
        for path, e in m.items():
            if not isinstance(e, DTensorEntry) or is_fully_replicated_entry(e) or rank in covered[path]:
                continue
            if path not in groups:
                groups[path] = replica_groups(e)
            for g in groups[path]:
                if rank in g:
                    covered[path] |= g
                    break
            parts[path].append(e)
    return {
        p: DTensorEntry(shards=sorted((s for e in es for s in e.shards), key=lambda sh: sh.offsets), mesh=es[0].mesh, dim_map=es[0].dim_map)
        for p, es in parts.items()
    }


def _remove_entry(manifest: Manifest, logical_path: str) -> None:
    if logical_path not in manifest:
        return
    del manifest[logical_path]
    parent_path, _, key = logical_path.rpartition("/")
    if not parent_path:
        return
    parent = manifest[parent_path]
    if is_dict_entry(parent):
        if key in parent.keys:
            parent.keys.remove(key)
        else:
            parent.keys.remove(int(key))


def get_manifest_for_rank(metadata: SnapshotMetadata, rank: int) -> Tuple[Manifest, Dict[str, Entry]]:
    per_rank = _split_by_rank(metadata)
    merged = _merge_sharded(per_rank)
    merged.update(_merge_dtensors(per_rank))
    if rank < metadata.world_size:
        local = dict(per_rank[rank])
        for path, e in per_rank[0].items():
            if is_fully_replicated_entry(e):
                local[path] = e
        for path, e in local.items():
            if isinstance(e, (ShardedTensorEntry, DTensorEntry)):
                # fully replicated DTensors are not merged (they have a single shard set); the reference
                # indexes `merged` unconditionally here (T:manifest_ops.py:82-84) and raises KeyError for them
                local[path] = merged.get(path, e)
        return local, merged
    # a rank that did not exist at save time: rank 0's view minus everything that is not replicated
    local = dict(per_rank[0])
    for path in list(local):
        e = local.get(path)
        if e is None or is_container_entry(e) or is_fully_replicated_entry(e):
            continue
        _remove_entry(local, path)
    return local, merged


def handle_sharded_tensor_elasticity(manifest: Manifest, merged_sd_entries: Dict[str, Entry], tensor_requests: List[str]) -> None:
    """Make the presence of sharded entries follow the *target* state dict (T:manifest_ops.py:180-247):
    requested-but-absent entries are added (a rank loads a table it did not save), present-but-unrequested
    ones are dropped."""
    if is_sharded_tensor_elasticity_enabled_at_root_only() and not all(len(p.split("/")) == 2 for p in merged_sd_entries):
        return
    wanted = [p for p in tensor_requests if p in merged_sd_entries]
    for path in wanted:
        if path not in manifest:
            manifest[path] = merged_sd_entries[path]
            parent, _, key = path.rpartition("/")
            manifest[parent].keys.append(key)
    for path in list(manifest):
        e = manifest[path]
        # fully replicated DTensors are ordinary replicated state, not elastic shards
        if isinstance(e, (ShardedTensorEntry, DTensorEntry)) and not is_fully_replicated_entry(e) and path not in wanted:
            del manifest[path]
