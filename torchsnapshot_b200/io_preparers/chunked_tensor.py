"""Tensors above the chunk-size knob are persisted as several dim-0 chunk files
(T:io_preparers/chunked_tensor.py:36-128).  The chunk plan is pure index arithmetic and must match the
reference's ``torch.chunk`` semantics exactly: n = ceil(bytes / limit) requested chunks, each
ceil(dim0 / n) rows, the last one ragged."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import torch

from ..io_types import Future, ReadReq, WriteReq
from ..knobs import get_max_chunk_size_bytes
from ..manifest import ChunkedTensorEntry, Shard
from ..serialization import dtype_to_string
from .tensor import PrepareFunc, TensorIOPreparer


@dataclass
class Chunk:
    offsets: List[int]
    sizes: List[int]
    dtype: str


def box_view(tensor: torch.Tensor, offsets: List[int], sizes: List[int]) -> torch.Tensor:
    """The sub-box [offsets, offsets+sizes) of `tensor` as a view (0-d tensors count as 1-d)."""
    view = tensor.reshape(1) if tensor.dim() == 0 else tensor
    for d, (o, s) in enumerate(zip(offsets, sizes)):
        if o != 0 or s != view.shape[d]:
            view = view.narrow(d, o, s)
    return view


class ChunkedTensorIOPreparer:
    @staticmethod
    def chunk_tensor(tensor: torch.Tensor, chunking_dim: int = 0, chunk_sz_bytes: Optional[int] = None) -> List[Chunk]:
        limit = chunk_sz_bytes or get_max_chunk_size_bytes()
        shape = [1] if tensor.dim() == 0 else list(tensor.shape)
        nbytes = tensor.numel() * tensor.element_size()
        wanted = math.ceil(nbytes / limit)
        extent = shape[chunking_dim]
        rows = math.ceil(extent / wanted) if wanted > 0 else extent
        plan: List[Chunk] = []
        lo = 0
        while lo < extent:
            n = min(rows, extent - lo)
            offsets = [0] * len(shape)
            offsets[chunking_dim] = lo
            sizes = list(shape)
            sizes[chunking_dim] = n
            plan.append(Chunk(offsets=offsets, sizes=sizes, dtype=str(tensor.dtype)))
            lo += n
        return plan

    @staticmethod
    def _get_subtensor_view(tensor: torch.Tensor, chunk: Union[Shard, Chunk]) -> torch.Tensor:
        return box_view(tensor, chunk.offsets, chunk.sizes)

    @classmethod
    def prepare_write(
        cls,
        storage_path: str,
        tensor: torch.Tensor,
        chunking_instruction: List[Chunk],
        is_async_snapshot: bool = False,
        _tensor_prepare_func: Optional[PrepareFunc] = None,
    ) -> Tuple[ChunkedTensorEntry, List[WriteReq]]:
        shards: List[Shard] = []
        reqs: List[WriteReq] = []
        for c in chunking_instruction:
            tag = "_".join(str(o) for o in c.offsets)
            e, wr = TensorIOPreparer.prepare_write(
                f"{storage_path}_{tag}", box_view(tensor, c.offsets, c.sizes), is_async_snapshot, _tensor_prepare_func
            )
            shards.append(Shard(offsets=c.offsets, sizes=c.sizes, tensor=e))
            reqs.extend(wr)
        entry = ChunkedTensorEntry(dtype=dtype_to_string(tensor.dtype), shape=list(tensor.shape), chunks=shards, replicated=False)
        return entry, reqs

    @classmethod
    def prepare_read(
        cls,
        entry: ChunkedTensorEntry,
        tensor_out: Optional[torch.Tensor] = None,
        buffer_size_limit_bytes: Optional[int] = None,
    ) -> Tuple[List[ReadReq], Future[torch.Tensor]]:
        if tensor_out is None or not TensorIOPreparer.can_load_inplace(entry, tensor_out):
            tensor_out = TensorIOPreparer.empty_tensor_from_entry(entry)
        reqs: List[ReadReq] = []
        for c in entry.chunks:
            rr, _ = TensorIOPreparer.prepare_read(c.tensor, box_view(tensor_out, c.offsets, c.sizes), buffer_size_limit_bytes)
            reqs.extend(rr)
        return reqs, Future(obj=tensor_out)
