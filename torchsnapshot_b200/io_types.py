"""The seam between planning and execution — same names and call contracts as T:io_types.py:24-120, so
third-party StoragePlugins and stagers written against the reference plug in unchanged."""
from __future__ import annotations

import abc
import asyncio
import io
from concurrent.futures import Executor
from dataclasses import dataclass, field
from typing import Generic, Optional, Tuple, TypeVar, Union

BufferType = Union[bytes, memoryview]
T = TypeVar("T")


class BufferStager(abc.ABC):
    """Produces the bytes of one storage object."""

    @abc.abstractmethod
    async def stage_buffer(self, executor: Optional[Executor] = None) -> BufferType: ...

    @abc.abstractmethod
    def get_staging_cost_bytes(self) -> int: ...


class BufferConsumer(abc.ABC):
    """Consumes the bytes of one read (whole object or byte range)."""

    @abc.abstractmethod
    async def consume_buffer(self, buf: bytes, executor: Optional[Executor] = None) -> None: ...

    @abc.abstractmethod
    def get_consuming_cost_bytes(self) -> int: ...


@dataclass
class WriteReq:
    path: str
    buffer_stager: BufferStager


@dataclass
class ReadReq:
    path: str
    buffer_consumer: BufferConsumer
    byte_range: Optional[Tuple[int, int]] = None


@dataclass
class Future(Generic[T]):
    obj: Optional[T] = None


@dataclass
class WriteIO:
    path: str
    buf: BufferType


@dataclass
class ReadIO:
    path: str
    buf: io.BytesIO = field(default_factory=io.BytesIO)
    byte_range: Optional[Tuple[int, int]] = None


def _loop(event_loop: Optional[asyncio.AbstractEventLoop]) -> asyncio.AbstractEventLoop:
    return event_loop if event_loop is not None else asyncio.new_event_loop()


class StoragePlugin(abc.ABC):
    @abc.abstractmethod
    async def write(self, write_io: WriteIO) -> None: ...

    @abc.abstractmethod
    async def read(self, read_io: ReadIO) -> None: ...

    @abc.abstractmethod
    async def delete(self, path: str) -> None: ...

    @abc.abstractmethod
    async def delete_dir(self, path: str) -> None: ...

    @abc.abstractmethod
    async def close(self) -> None: ...

    def sync_write(self, write_io: WriteIO, event_loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
        _loop(event_loop).run_until_complete(self.write(write_io))

    def sync_read(self, read_io: ReadIO, event_loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
        _loop(event_loop).run_until_complete(self.read(read_io))

    def sync_close(self, event_loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
        _loop(event_loop).run_until_complete(self.close())
