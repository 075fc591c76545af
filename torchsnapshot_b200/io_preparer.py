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


def prepare_write(
    obj: Any,
    logical_path: str,
    rank: int,
    replicated: bool,
    is_async_snapshot: bool = False,
    _tensor_prepare_func: Optional[PrepareFunc] = None,
) -> Tuple[Entry, List[WriteReq]]:
    if PrimitivePreparer.should_inline(obj):
        entry = PrimitivePreparer.prepare_write(obj)
        entry.replicated = replicated
        return entry, []
    path = get_storage_path(obj, logical_path, rank, replicated)
    if isinstance(obj, ShardedTensor):
        return ShardedTensorIOPreparer.prepare_write(path, obj, is_async_snapshot, _tensor_prepare_func)
    if isinstance(obj, DTensor):
        return DTensorIOPreparer.prepare_write(path, obj, is_async_snapshot, _tensor_prepare_func)
    if isinstance(obj, torch.Tensor):
        if obj.numel() * obj.element_size() > get_max_chunk_size_bytes():
            plan = ChunkedTensorIOPreparer.chunk_tensor(obj)
            entry, reqs = ChunkedTensorIOPreparer.prepare_write(path, obj, plan, is_async_snapshot, _tensor_prepare_func)
        else:
            entry, reqs = TensorIOPreparer.prepare_write(path, obj, is_async_snapshot, _tensor_prepare_func)
    else:
        entry, reqs = ObjectIOPreparer.prepare_write(path, obj)
    entry.replicated = replicated
    return entry, reqs


def prepare_read(
    entry: Entry, obj_out: Optional[Any] = None, buffer_size_limit_bytes: Optional[int] = None
) -> Tuple[List[ReadReq], Future[Any]]:
    if isinstance(entry, ShardedTensorEntry):
        return ShardedTensorIOPreparer.prepare_read(entry, obj_out)
    if isinstance(entry, ChunkedTensorEntry):
        return ChunkedTensorIOPreparer.prepare_read(entry, obj_out, buffer_size_limit_bytes=buffer_size_limit_bytes)
    if isinstance(entry, DTensorEntry):
        return DTensorIOPreparer.prepare_read(entry, obj_out)
    if isinstance(entry, TensorEntry):
        return TensorIOPreparer.prepare_read(entry, obj_out, buffer_size_limit_bytes=buffer_size_limit_bytes)
    if isinstance(entry, ObjectEntry):
        return ObjectIOPreparer.prepare_read(entry, obj_out)
    if isinstance(entry, PrimitiveEntry):
        return PrimitivePreparer.prepare_read(entry)
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
