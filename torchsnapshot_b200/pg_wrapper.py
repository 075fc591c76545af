"""Thin process-group adapter (T:pg_wrapper.py:17-91): object collectives that degrade to no-ops when
torch.distributed is not initialised.  Control plane only — KB-sized pickles, latency bound."""
from __future__ import annotations

from typing import Any, List, Optional

import torch
import torch.distributed as dist


# number of collectives issued through PGWrapper objects of this process (bench.py reports the per-take count)
COLLECTIVE_COUNT = {"total": 0}


def _count() -> None:
    COLLECTIVE_COUNT["total"] += 1


class PGWrapper:
    def __init__(self, pg: Optional[dist.ProcessGroup] = None) -> None:
        self.pg = pg if pg is not None else (dist.group.WORLD if dist.is_available() and dist.is_initialized() else None)

    def get_rank(self) -> int:
        return 0 if self.pg is None else dist.get_rank(group=self.pg)

    def get_world_size(self) -> int:
        return 1 if self.pg is None else dist.get_world_size(group=self.pg)

    def _alone(self) -> bool:
        # a one-rank group needs no communication: object collectives would still pickle and copy (≈0.3–2 ms each)
        return self.pg is None or self.get_world_size() == 1

    def barrier(self) -> None:
        if not self._alone():
            _count()
            if dist.get_backend(self.pg) == "nccl":
                dist.barrier(group=self.pg, device_ids=[torch.cuda.current_device()])
            else:
                dist.barrier(group=self.pg)

    def broadcast_object_list(self, obj_list: List[Any], src: int = 0) -> None:
        if not self._alone():
            _count()
            dist.broadcast_object_list(obj_list, src=dist.get_global_rank(self.pg, src), group=self.pg)

    def all_gather_object(self, obj_list: List[Any], obj: Any) -> None:
        if self._alone():
            obj_list[0] = obj
        else:
            _count()
            dist.all_gather_object(obj_list, obj, group=self.pg)

    def scatter_object_list(self, output_list: List[Any], input_list: Optional[List[Any]], src: int = 0) -> None:
        """output_list[0] <- input_list[rank] of rank `src` (T:pg_wrapper.py:60-91; NCCL has no object scatter, so
        it is emulated with a broadcast there)."""
        rank, world = self.get_rank(), self.get_world_size()
        if rank == src:
            if input_list is None:
                raise RuntimeError("The src rank's input_list for scatter_object_list must not be None.")
            if len(input_list) != world:
                raise RuntimeError(
                    f"The length of input_list {len(input_list)} for scatter_object_list "
                    f"must be the same as the process group's world size ({world})."
                )
        if self._alone():
            output_list[0] = input_list[0]  # type: ignore[index]
            return
        if dist.get_backend(self.pg) == "nccl":
            payload = list(input_list) if rank == src else [None] * world
            self.broadcast_object_list(payload, src=src)
            output_list[0] = payload[rank]
            return
        _count()
        dist.scatter_object_list(output_list, input_list if rank == src else None, src=dist.get_global_rank(self.pg, src), group=self.pg)
