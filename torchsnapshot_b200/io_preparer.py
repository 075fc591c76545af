"""Type dispatch from a flattened leaf to its preparer, and the storage-path scheme
(T:io_preparer.py:52-182): ``replicated_sharded/ | sharded/ | replicated/ | <rank>/`` + logical path."""
from __future__ import annotations

import os
from typing import Any, List, Optional, Tuple

import torch
from torch.distributed._shard.sharded_tensor import ShardedTensor
from torch.distributed.tensor import DTensor, Shard as ShardPlacement

from .io_preparers.chunked_tensor import Chunk, ChunkedTensorIOPreparer
from .io_preparers.dtensor import DTensorIOPreparer
from .io_preparers.object import ObjectBufferConsumer, ObjectBufferStager, ObjectIOPreparer
from .io_preparers.sharded_tensor import ShardedTensorBufferConsumer, ShardedTensorIOPreparer
from .io_preparers.tensor import PrepareFunc, TensorBufferConsumer, TensorBufferStager, TensorIOPreparer, tensor_copy
from .io_types import Future, ReadReq, WriteReq
from .knobs import get_max_chunk_size_bytes
from .manifest import (
    ChunkedTensorEntry,
    DTensorEntry,
    Entry,
    ObjectEntry,
    PrimitiveEntry,
    ShardedTensorEntry,
    TensorEntry,
)


def is_sharded(obj: Any) -> bool:
    if isinstance(obj, ShardedTensor):
        return True
    if isinstance(obj, DTensor):
        return any(isinstance(p, ShardPlacement) for p in obj.placements)
    return False


def get_storage_path(obj: Any, logical_path: str, rank: int, replicated: bool) -> str:
    if is_sharded(obj):
        root = "replicated_sharded" if replicated else "sharded"
    else:
        root = "replicated" if replicated else str(rank)
    return os.path.join(root, logical_path)


class PrimitivePreparer:
    @staticmethod
    def should_inline(obj: Any) -> bool:
        return type(obj).__name__ in PrimitiveEntry.supported_types

    @staticmethod
    def prepare_write(obj: Any) -> PrimitiveEntry:
        return PrimitiveEntry.from_object(obj)

    @staticmethod
    def prepare_read(entry: PrimitiveEntry) -> Tuple[List[ReadReq], Future[Any]]:
        return [], Future(obj=entry.get_value())


def _tensor_like_writer(obj: Any):
    """Preparer for tensor-like leaves, None for anything else."""
    if isinstance(obj, ShardedTensor):
        return ShardedTensorIOPreparer.prepare_write
    if isinstance(obj, DTensor):
        return DTensorIOPreparer.prepare_write
    return None


def prepare_write(
    obj: Any,
    logical_path: str,
    rank: int,
    replicated: bool,
    is_async_snapshot: bool = False,
    _tensor_prepare_func: Optional[PrepareFunc] = None,
) -> Tuple[Entry, List[WriteReq]]:
    """Leaf -> (manifest entry, write requests).  Primitives are inlined in the metadata; sharded leaves keep
    their own replication bookkeeping (per-shard entries), everything else is flagged here (T:io_preparer.py:82-147)."""
    if PrimitivePreparer.should_inline(obj):
        inline = PrimitivePreparer.prepare_write(obj)
        inline.replicated = replicated
        return inline, []
    path = get_storage_path(obj, logical_path, rank, replicated)
    sharded_writer = _tensor_like_writer(obj)
    if sharded_writer is not None:
        return sharded_writer(path, obj, is_async_snapshot, _tensor_prepare_func)
    if not isinstance(obj, torch.Tensor):
        entry, reqs = ObjectIOPreparer.prepare_write(path, obj)
    elif obj.numel() * obj.element_size() > get_max_chunk_size_bytes():
        entry, reqs = ChunkedTensorIOPreparer.prepare_write(
            path, obj, ChunkedTensorIOPreparer.chunk_tensor(obj), is_async_snapshot, _tensor_prepare_func
        )
    else:
        entry, reqs = TensorIOPreparer.prepare_write(path, obj, is_async_snapshot, _tensor_prepare_func)
    entry.replicated = replicated
    return entry, reqs


def prepare_read(
    entry: Entry, obj_out: Optional[Any] = None, buffer_size_limit_bytes: Optional[int] = None
) -> Tuple[List[ReadReq], Future[Any]]:
    """Manifest entry (+ optional in-place target) -> (read requests, future).  T:io_preparer.py:150-182."""
    budgeted = {"buffer_size_limit_bytes": buffer_size_limit_bytes}
    readers = (
        (ShardedTensorEntry, lambda: ShardedTensorIOPreparer.prepare_read(entry, obj_out)),
        (ChunkedTensorEntry, lambda: ChunkedTensorIOPreparer.prepare_read(entry, obj_out, **budgeted)),
        (DTensorEntry, lambda: DTensorIOPreparer.prepare_read(entry, obj_out)),
        (TensorEntry, lambda: TensorIOPreparer.prepare_read(entry, obj_out, **budgeted)),
        (ObjectEntry, lambda: ObjectIOPreparer.prepare_read(entry, obj_out)),
        (PrimitiveEntry, lambda: PrimitivePreparer.prepare_read(entry)),
    )
    for kind, make in readers:
        if isinstance(entry, kind):
            return make()
    raise Exception(f"Unsupported entry type: {entry} ({entry.type}).")


__all__ = [
    "Chunk",
    "ObjectBufferConsumer",
    "ObjectBufferStager",
    "ShardedTensorBufferConsumer",
    "TensorBufferConsumer",
    "TensorBufferStager",
    "tensor_copy",
    "prepare_read",
    "prepare_write",
    "get_storage_path",
    "is_sharded",
]
