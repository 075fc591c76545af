"""DTensor leg (FSDP2 / HSDP layouts).  Same shape as the ShardedTensor leg: the local shard's
global box comes from ``compute_local_shape_and_global_offset``; it is subdivided along the largest
sharded dim (T:io_preparers/dtensor.py:64-98, 123-198) and read back through box intersection
(200-278)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch.distributed.tensor import DTensor, Replicate, Shard as ShardPlacement
from torch.distributed.tensor._utils import compute_local_shape_and_global_offset

from ..io_types import Future, ReadReq, WriteReq
from ..knobs import get_max_shard_size_bytes
from ..manifest import DTensorEntry, Shard
from .sharded_tensor import ShardedTensorIOPreparer, overlap_read_reqs
from .tensor import PrepareFunc, TensorIOPreparer


class DTensorIOPreparer:
    @staticmethod
    def _get_largest_shard_dim(local_shape: Sequence[int], mesh, placements) -> int:
        dims = [p.dim for p in placements if isinstance(p, ShardPlacement)]
        if dims:
            return max(dims, key=lambda d: local_shape[d])
        # fully replicated: subdivide along the longest dim
        return max(range(len(local_shape)), key=lambda d: local_shape[d]) if len(local_shape) else 0

    @staticmethod
    def _get_dim_map(obj: DTensor) -> List[List[int]]:
        dim_map: List[List[int]] = [[] for _ in range(obj.dim())]
        for mesh_dim, p in enumerate(obj.placements):
            if isinstance(p, ShardPlacement):
                dim_map[p.dim].append(mesh_dim)
            elif not isinstance(p, Replicate):
                raise ValueError("Unsupported placement type")
        return [d if d else [-1] for d in dim_map]

    @staticmethod
    def _get_global_shape(entry: DTensorEntry) -> List[int]:
        shape = [0] * len(entry.shards[0].sizes)
        for sh in entry.shards:
            for d, (o, s) in enumerate(zip(sh.offsets, sh.sizes)):
                shape[d] = max(shape[d], o + s)
        return shape

    @classmethod
    def prepare_write(
        cls,
        storage_path: str,
        obj: DTensor,
        is_async_snapshot: bool = False,
        _tensor_prepare_func: Optional[PrepareFunc] = None,
    ) -> Tuple[DTensorEntry, List[WriteReq]]:
        mesh, placements = obj.device_mesh, obj.placements
        local_shape, offsets = compute_local_shape_and_global_offset(obj.size(), mesh, placements)
        dim = cls._get_largest_shard_dim(local_shape, mesh, placements)
        shards: List[Shard] = []
        reqs: List[WriteReq] = []
        for view, off, sz in ShardedTensorIOPreparer.subdivide_shard(
            obj.to_local(), list(offsets), list(local_shape), dim, get_max_shard_size_bytes()
        ):
            tag = "_".join(str(i) for i in off)
            e, wr = TensorIOPreparer.prepare_write(f"{storage_path}_{tag}", view, is_async_snapshot, _tensor_prepare_func)
            reqs.extend(wr)
            shards.append(Shard(offsets=off, sizes=sz, tensor=e))
        entry = DTensorEntry(shards=shards, mesh=mesh.mesh.cpu().numpy().tolist(), dim_map=cls._get_dim_map(obj))
        return entry, reqs

    @classmethod
    def prepare_read(cls, entry: DTensorEntry, obj_out: Optional[DTensor] = None) -> Tuple[List[ReadReq], Future[DTensor]]:
        if obj_out is None:
            raise RuntimeError("No output DTensor object found. Cannot read a DTensorEntry without a runtime object.")
        local_shape, offsets = compute_local_shape_and_global_offset(obj_out.shape, obj_out.device_mesh, obj_out.placements)
        local = obj_out.to_local()
        assert tuple(local.shape) == tuple(local_shape)
        return overlap_read_reqs(entry.shards, [(local, list(offsets), list(local_shape))]), Future(obj=obj_out)
