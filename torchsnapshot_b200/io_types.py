"""The seam between planning and execution.

Names and call contracts are those of the reference's ``io_types`` module (T:io_types.py:24-120) because this IS
the drop-in boundary on the Python side: third-party storage plugins and stagers written against torchsnapshot
plug in unchanged, and the engine's scheduler recognises the reference's own objects by these attribute names.

Contract summary
    BufferStager.stage_buffer(executor)   -> bytes-like of exactly the storage object's size; called on the event
                                             loop; the returned buffer is owned by the pipeline until written.
    BufferStager.get_staging_cost_bytes() -> upper bound of host bytes held while staging (budget accounting).
    BufferConsumer.consume_buffer(buf, executor) <- the bytes of ``ReadReq.byte_range`` (or the whole object).
    StoragePlugin.{write,read,delete,delete_dir,close} are coroutines; paths are relative to the snapshot root;
    ``ReadIO.byte_range`` is half-open ``[lo, hi)``.
"""
from __future__ import annotations

import abc
import asyncio
import io
from concurrent.futures import Executor
from dataclasses import dataclass, field
from typing import Generic, Optional, Tuple, TypeVar, Union

BufferType = Union[bytes, memoryview]
T = TypeVar("T")
ByteRange = Tuple[int, int]


class BufferStager(abc.ABC):
    """Produces the bytes of one storage object (one file / one slab)."""

    @abc.abstractmethod
    async def stage_buffer(self, executor: Optional[Executor] = None) -> BufferType:
        raise NotImplementedError

    @abc.abstractmethod
    def get_staging_cost_bytes(self) -> int:
        raise NotImplementedError


class BufferConsumer(abc.ABC):
    """Consumes the bytes of one read."""

    @abc.abstractmethod
    async def consume_buffer(self, buf: bytes, executor: Optional[Executor] = None) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    def get_consuming_cost_bytes(self) -> int:
        raise NotImplementedError


@dataclass
class WriteReq:
    """``path`` (relative to the snapshot root) receives what ``buffer_stager`` produces."""

    path: str
    buffer_stager: BufferStager


@dataclass
class ReadReq:
    """``buffer_consumer`` receives ``path[byte_range]`` (the whole object when the range is None)."""

    path: str
    buffer_consumer: BufferConsumer
    byte_range: Optional[ByteRange] = None


@dataclass
class Future(Generic[T]):
    """Where a read deposits its result; for in-place loads ``obj`` is set at planning time."""

    obj: Optional[T] = None


@dataclass
class WriteIO:
    path: str
    buf: BufferType


@dataclass
class ReadIO:
    path: str
    buf: io.BytesIO = field(default_factory=io.BytesIO)
    byte_range: Optional[ByteRange] = None


class StoragePlugin(abc.ABC):
    """Asynchronous object store rooted at a snapshot directory / prefix."""

    @abc.abstractmethod
    async def write(self, write_io: WriteIO) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    async def read(self, read_io: ReadIO) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    async def delete(self, path: str) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    async def delete_dir(self, path: str) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    async def close(self) -> None:
        raise NotImplementedError

    # blocking conveniences used for the metadata file
    def _run(self, coro, event_loop: Optional[asyncio.AbstractEventLoop]) -> None:
        loop = event_loop if event_loop is not None else asyncio.new_event_loop()
        loop.run_until_complete(coro)

    def sync_write(self, write_io: WriteIO, event_loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
        self._run(self.write(write_io), event_loop)

    def sync_read(self, read_io: ReadIO, event_loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
        self._run(self.read(read_io), event_loop)

    def sync_close(self, event_loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
        self._run(self.close(), event_loop)
