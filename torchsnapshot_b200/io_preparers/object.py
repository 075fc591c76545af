"""Non-tensor leaves (tuples, optimizer hyper-parameters, ...) keep the reference's ``torch.save``
representation (T:io_preparers/object.py:33-95).  No bulk data: out of scope of the engine."""
from __future__ import annotations

import sys
from concurrent.futures import Executor
from typing import Any, Generic, List, Optional, Tuple, TypeVar

from ..io_types import BufferConsumer, BufferStager, BufferType, Future, ReadReq, WriteReq
from ..manifest import ObjectEntry
from ..serialization import Serializer, torch_load_from_bytes, torch_save_as_bytes

T = TypeVar("T")


class ObjectBufferStager(BufferStager):
    def __init__(self, obj: Any) -> None:
        self.obj = obj

    async def stage_buffer(self, executor: Optional[Executor] = None) -> BufferType:
        return torch_save_as_bytes(self.obj)

    def get_staging_cost_bytes(self) -> int:
        return sys.getsizeof(self.obj)


class ObjectBufferConsumer(BufferConsumer, Generic[T]):
    def __init__(self, fut: Future[T]) -> None:
        self.fut = fut
        self.cost = sys.getsizeof(fut.obj)

    async def consume_buffer(self, buf: bytes, executor: Optional[Executor] = None) -> None:
        self.fut.obj = torch_load_from_bytes(buf)

    def get_consuming_cost_bytes(self) -> int:
        return self.cost


class ObjectIOPreparer(Generic[T]):
    @staticmethod
    def prepare_write(storage_path: str, obj: T) -> Tuple[ObjectEntry, List[WriteReq]]:
        kind = f"{type(obj).__module__}.{type(obj).__name__}"
        entry = ObjectEntry(location=storage_path, serializer=Serializer.TORCH_SAVE.value, obj_type=kind, replicated=False)
        return entry, [WriteReq(path=storage_path, buffer_stager=ObjectBufferStager(obj))]

    @classmethod
    def prepare_read(cls, entry: ObjectEntry, obj_out: Optional[Any] = None) -> Tuple[List[ReadReq], Future[T]]:
        fut: Future[T] = Future(obj=obj_out)
        return [ReadReq(path=entry.location, buffer_consumer=ObjectBufferConsumer(fut))], fut
