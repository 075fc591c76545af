"""Execution of write/read plans.

The reference drives every request through an asyncio state machine (stage -> write, budget gated,
4 staging threads, <=16 I/Os; T:scheduler.py:222-339, 386-446).  Here requests are split in two:

  * raw tensor traffic against a local filesystem — the hot path — becomes ONE engine job per device:
    all members are described to the C ABI, packed by a single kernel launch pair, drained through
    the pinned ring and written by native workers (save), or read, uploaded and scattered (restore);
  * everything else (pickled leaves, third-party storage plugins) keeps the reference's pipeline
    semantics: staging is admitted while the host-memory budget allows, at most
    ``get_max_per_rank_io_concurrency()`` storage operations are in flight.

``execute_write_reqs`` returns when every source tensor has been read (the async_take gate,
T:scheduler.py:299); ``PendingIOWork.complete`` drains the I/O."""
from __future__ import annotations

import asyncio
import logging
import os
import socket
import time
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Set, Tuple

import psutil
import torch

from . import _native
from .io_types import ReadIO, ReadReq, StoragePlugin, WriteIO, WriteReq
from .knobs import get_max_per_rank_io_concurrency, get_memory_budget_override
from .native_plan import describe_consumer, describe_stager, native_root as _native_root
from .pg_wrapper import PGWrapper

logger = logging.getLogger(__name__)

_MAX_PER_RANK_MEMORY_BUDGET_BYTES = 32 * 1024**3
_AVAILABLE_MEMORY_MULTIPLIER = 0.6
_MAX_PER_RANK_CPU_CONCURRENCY = 4


_LOCAL_WORLD: Dict[int, Tuple[object, int]] = {}  # id(pg) -> (pg kept alive so the id stays unique, ranks on this host)


def get_local_world_size(pg: PGWrapper) -> int:
    """Ranks of `pg` on this host (T:scheduler.py:35-44).  Host placement does not change during a job, so the
    hostname all-gather is paid once per process group instead of once per snapshot."""
    key = id(pg.pg)
    ent = _LOCAL_WORLD.get(key)
    if ent is None or ent[0] is not pg.pg:
        me = socket.gethostname()
        names = [None] * pg.get_world_size()
        pg.all_gather_object(names, me)
        ent = (pg.pg, sum(1 for n in names if n == me))
        _LOCAL_WORLD[key] = ent
    return ent[1]


def get_process_memory_budget_bytes(pg: PGWrapper) -> int:
    override = get_memory_budget_override()
    if override is not None:
        return override
    avail = int(psutil.virtual_memory().available * _AVAILABLE_MEMORY_MULTIPLIER)
    return min(avail // get_local_world_size(pg), _MAX_PER_RANK_MEMORY_BUDGET_BYTES)


# ---- routing -------------------------------------------------------------------------------------
def _engine_key(tensors: List[torch.Tensor]) -> int:
    for t in tensors:
        if t.is_cuda:
            return t.device.index if t.device.index is not None else torch.cuda.current_device()
    return -1


class _NativeJobs:
    """One engine job per device that appears in the plan."""

    def __init__(self, save: bool) -> None:
        self.save = save
        self.jobs: Dict[int, "_native.Job"] = {}
        self.payload_bytes = 0

    def job_for(self, key: int) -> "_native.Job":
        j = self.jobs.get(key)
        if j is None:
            eng = _native.get_engine(key)
            j = eng.save_job() if self.save else eng.load_job()
            self.jobs[key] = j
        return j

    def submit(self) -> None:
        for key, j in self.jobs.items():
            stream = torch.cuda.current_stream(key).cuda_stream if key >= 0 else None
            j.submit(stream)

    def wait_device(self) -> None:
        # failures (e.g. an unwritable path) are reported by wait(), i.e. by PendingSnapshot.wait() for
        # async_take — the same place the reference surfaces storage errors (tests/test_async_take.py:58-66)
        for j in self.jobs.values():
            try:
                j.wait_device()
            except _native.NativeError:
                pass

    def wait(self) -> None:
        err = None
        for j in self.jobs.values():
            try:
                j.wait()
            except Exception as e:  # keep draining the others before raising
                err = err or e
        if err is not None:
            raise err

    def stats(self) -> List[dict]:
        return [j.stats() for j in self.jobs.values()]

    def traces(self) -> List[List[dict]]:
        """Per-chunk timelines (only when the engine was created with ENGINE_TRACE)."""
        return [j.trace() for j in self.jobs.values() if j.engine.flags & _native.ENGINE_TRACE]

    def destroy(self) -> None:
        for j in self.jobs.values():
            j.destroy()
        self.jobs.clear()


LAST_STATS: Dict[str, object] = {}  # filled by the last executed plan; read by bench.py


class PendingIOWork:
    def __init__(self, native: Optional[_NativeJobs], io_tasks: Set["asyncio.Task"], begin_ts: float, rank: int, nbytes: int) -> None:
        self.native = native
        self.io_tasks = io_tasks
        self.begin_ts = begin_ts
        self.rank = rank
        self.nbytes = nbytes

    async def complete(self) -> None:
        loop = asyncio.get_running_loop()
        try:
            if self.io_tasks:
                await asyncio.gather(*self.io_tasks)
            if self.native is not None:
                await loop.run_in_executor(None, self.native.wait)
                LAST_STATS["save"] = self.native.stats()
                LAST_STATS["save_trace"] = self.native.traces()
        finally:
            if self.native is not None:
                self.native.destroy()
        dt = max(time.monotonic() - self.begin_ts, 1e-9)
        logger.info(f"Rank {self.rank} completed writing in {dt:.2f} seconds (throughput {self.nbytes / 2**20 / dt:.2f}MB/s)")

    def sync_complete(self, event_loop: asyncio.AbstractEventLoop) -> None:
        event_loop.run_until_complete(self.complete())


async def execute_write_reqs(
    write_reqs: List[WriteReq], storage: StoragePlugin, memory_budget_bytes: int, rank: int
) -> PendingIOWork:
    begin = time.monotonic()
    phases: Dict[str, float] = {}
    t_mark = [time.perf_counter()]

    def lap(name: str) -> None:
        now = time.perf_counter()
        phases[name] = (now - t_mark[0]) * 1e3
        t_mark[0] = now

    loop = asyncio.get_running_loop()
    root = _native_root(storage, "write")
    native: Optional[_NativeJobs] = None
    generic: List[WriteReq] = []
    total = 0
    for wr in write_reqs:
        described = describe_stager(wr.buffer_stager) if root is not None else None
        if described is not None:
            descs, keep, nbytes = described
            if native is None:
                native = _NativeJobs(save=True)
            job = native.job_for(_engine_key(keep))
            fi = job.add_file(os.path.join(root, wr.path), nbytes)
            for d in descs:
                job.add_member(fi, d)
            job._keepalive.extend(keep)
            total += nbytes
        else:
            generic.append(wr)
    lap("describe+build_job")
    if native is not None:
        native.submit()
    lap("submit")

    # generic pipeline: budget-gated staging, bounded concurrent writes
    io_tasks: Set[asyncio.Task] = set()
    if generic:
        executor = ThreadPoolExecutor(max_workers=_MAX_PER_RANK_CPU_CONCURRENCY)
        budget = memory_budget_bytes
        io_slots = asyncio.Semaphore(get_max_per_rank_io_concurrency())
        credit = asyncio.Condition()
        inflight = 0

        async def write_out(wr: WriteReq, buf) -> int:
            nonlocal budget, inflight
            n = len(buf)
            try:
                async with io_slots:
                    await storage.write(WriteIO(path=wr.path, buf=buf))
            finally:
                del buf
                async with credit:
                    budget += n
                    inflight -= 1
                    credit.notify_all()
            return n

        pending = sorted(generic, key=lambda w: w.buffer_stager.get_staging_cost_bytes())
        staging: Set[asyncio.Task] = set()

        async def stage_one(wr: WriteReq, cost: int) -> None:
            nonlocal budget
            buf = await wr.buffer_stager.stage_buffer(executor)
            async with credit:
                budget += cost - len(buf)  # the estimate is replaced by the real footprint
            io_tasks.add(asyncio.ensure_future(write_out(wr, buf)))

        for wr in pending:
            cost = wr.buffer_stager.get_staging_cost_bytes()
            async with credit:
                # an over-budget request is only admitted when nothing else is in flight (T:scheduler.py:266-272)
                await credit.wait_for(lambda: cost < budget or inflight == 0)
                budget -= cost
                inflight += 1
            staging.add(asyncio.ensure_future(stage_one(wr, cost)))
            total += cost
        if staging:
            await asyncio.gather(*staging)
        executor.shutdown(wait=False)

    lap("generic_staging")
    if native is not None:
        await loop.run_in_executor(None, native.wait_device)
    lap("wait_device")
    LAST_STATS["write_phases_ms"] = phases
    logger.info(f"Rank {rank} completed staging in {time.monotonic() - begin:.2f} seconds")
    return PendingIOWork(native, io_tasks, begin, rank, total)


def sync_execute_write_reqs(
    write_reqs: List[WriteReq], storage: StoragePlugin, memory_budget_bytes: int, rank: int, event_loop: asyncio.AbstractEventLoop
) -> PendingIOWork:
    return event_loop.run_until_complete(execute_write_reqs(write_reqs, storage, memory_budget_bytes, rank))


async def execute_read_reqs(read_reqs: List[ReadReq], storage: StoragePlugin, memory_budget_bytes: int, rank: int) -> None:
    begin = time.monotonic()
    loop = asyncio.get_running_loop()
    root = _native_root(storage, "read")
    native: Optional[_NativeJobs] = None
    generic: List[ReadReq] = []
    total = 0
    for rr in read_reqs:
        described = describe_consumer(rr.buffer_consumer) if root is not None else None
        if described is not None:
            descs, keep, wire_nbytes = described
            if rr.byte_range is not None:
                lo, hi = rr.byte_range
            else:
                lo, hi = 0, wire_nbytes
            if not descs:
                continue
            if native is None:
                native = _NativeJobs(save=False)
            job = native.job_for(_engine_key(keep))
            fi = job.add_file(os.path.join(root, rr.path), hi - lo, offset=lo)
            for d in descs:
                job.add_member(fi, d)
            job._keepalive.extend(keep)
            total += hi - lo
        else:
            generic.append(rr)
    try:
        if native is not None:
            native.submit()
        if generic:
            executor = ThreadPoolExecutor(max_workers=_MAX_PER_RANK_CPU_CONCURRENCY)
            io_slots = asyncio.Semaphore(get_max_per_rank_io_concurrency())
            budget = memory_budget_bytes
            credit = asyncio.Condition()
            inflight = 0

            async def one(rr: ReadReq, cost: int) -> None:
                nonlocal budget, inflight
                try:
                    async with io_slots:
                        rio = ReadIO(path=rr.path, byte_range=rr.byte_range)
                        await storage.read(rio)
                    buf = rio.buf.getbuffer()
                    await rr.buffer_consumer.consume_buffer(buf, executor)
                finally:
                    async with credit:
                        budget += cost
                        inflight -= 1
                        credit.notify_all()

            tasks = []
            for rr in generic:
                cost = rr.buffer_consumer.get_consuming_cost_bytes()
                async with credit:
                    await credit.wait_for(lambda: cost < budget or inflight == 0)
                    budget -= cost
                    inflight += 1
                tasks.append(asyncio.ensure_future(one(rr, cost)))
                total += cost
            await asyncio.gather(*tasks)
            executor.shutdown(wait=False)
        if native is not None:
            await loop.run_in_executor(None, native.wait)
            LAST_STATS["load"] = native.stats()
            LAST_STATS["load_trace"] = native.traces()
    finally:
        if native is not None:
            native.destroy()
    dt = max(time.monotonic() - begin, 1e-9)
    logger.info(f"Rank {rank} finished loading. Throughput: {total / 2**20 / dt:.2f}MB/s")


def sync_execute_read_reqs(
    read_reqs: List[ReadReq], storage: StoragePlugin, memory_budget_bytes: int, rank: int, event_loop: asyncio.AbstractEventLoop
) -> None:
    event_loop.run_until_complete(execute_read_reqs(read_reqs, storage, memory_budget_bytes, rank))
