"""``Snapshot.take / async_take / restore / read_object`` — the reference's public surface
(T:snapshot.py:67-1068) over the B200 data plane.

Only the orchestration lives here: gather keys, flatten state dicts, plan (prepare_write -> partition ->
batch), hand the plan to the scheduler/engine, commit ``.snapshot_metadata`` last.  On-disk format and
collective call pattern follow the reference so that snapshots are interchangeable; the data plane
underneath is the engine (see scheduler.py)."""
from __future__ import annotations

import asyncio
import copy
import fnmatch
import functools
import itertools
import logging
import os
import socket
import sys
import threading
import traceback
from collections import defaultdict
from datetime import timedelta
from typing import Any, Callable, Dict, List, Optional, Set, Tuple, TypeVar

import torch
import torch.distributed as dist
from torch.distributed._shard.sharded_tensor import ShardedTensor
from torch.distributed.tensor import DTensor
from torch.nn.parallel import DistributedDataParallel as DDP

from .batcher import batch_read_requests, batch_write_requests
from .flatten import flatten, inflate
from .io_preparer import is_sharded, prepare_read, prepare_write
from .io_types import ReadIO, ReadReq, StoragePlugin, WriteIO, WriteReq
from .knobs import is_batching_disabled
from .manifest import (
    SNAPSHOT_FORMAT_VERSION,
    Entry,
    Manifest,
    PrimitiveEntry,
    SnapshotMetadata,
    is_container_entry,
)
from .manifest_ops import get_manifest_for_rank, handle_sharded_tensor_elasticity
from .partitioner import consolidate_replicated_entries, partition_write_reqs
from .pg_wrapper import PGWrapper
from .scheduler import (
    _MAX_PER_RANK_MEMORY_BUDGET_BYTES,
    PendingIOWork,
    get_process_memory_budget_bytes,
    sync_execute_read_reqs,
    sync_execute_write_reqs,
)
from .stateful import AppState, RNGState, Stateful
from .storage_plugin import url_to_storage_plugin_in_event_loop

logger = logging.getLogger(__name__)

SNAPSHOT_METADATA_FNAME = ".snapshot_metadata"
T = TypeVar("T")
CustomPrepareFunc = Callable[[str, torch.Tensor, bool], torch.Tensor]


class Snapshot:
    """A handle to a snapshot at ``path``; see :meth:`take`, :meth:`async_take`, :meth:`restore`."""

    def __init__(self, path: str, pg: Optional[dist.ProcessGroup] = None, storage_options: Optional[Dict[str, Any]] = None) -> None:
        self.path = path
        self.pg = pg
        self._storage_options = storage_options
        self._metadata: Optional[SnapshotMetadata] = None

    # ---- save ---------------------------------------------------------------------------------------
    @classmethod
    def take(
        cls,
        path: str,
        app_state: AppState,
        pg: Optional[dist.ProcessGroup] = None,
        replicated: Optional[List[str]] = None,
        storage_options: Optional[Dict[str, Any]] = None,
        _custom_tensor_prepare_func: Optional[CustomPrepareFunc] = None,
    ) -> "Snapshot":
        cls._validate_app_state(app_state)
        loop = asyncio.new_event_loop()
        pgw = PGWrapper(pg)
        path, globs, keys = cls._coalesce_path_and_replicated(path, pgw, app_state, replicated or [])
        storage = url_to_storage_plugin_in_event_loop(path, loop, storage_options)
        try:
            pending, metadata = cls._take_impl(path, app_state, globs, pgw, storage, loop, False, _custom_tensor_prepare_func, keys)
            # the engine is draining in its own threads: encode the metadata meanwhile (json with indent=2
            # runs in the pure-Python encoder, ~40 ms for a 300-entry manifest)
            encoded = metadata.to_yaml().encode("utf-8") if pgw.get_rank() == 0 else None
            pending.sync_complete(loop)
            # commit point: metadata goes last, after every rank finished writing (T:snapshot.py:202-209)
            pgw.barrier()
            if pgw.get_rank() == 0:
                cls._write_snapshot_metadata(metadata, storage, loop, encoded)
        finally:
            storage.sync_close(loop)
            loop.close()
        snap = cls(path=path, pg=pg, storage_options=storage_options)
        snap._metadata = metadata
        return snap

    @classmethod
    def async_take(
        cls,
        path: str,
        app_state: AppState,
        pg: Optional[dist.ProcessGroup] = None,
        replicated: Optional[List[str]] = None,
        storage_options: Optional[Dict[str, Any]] = None,
        _custom_tensor_prepare_func: Optional[CustomPrepareFunc] = None,
    ) -> "PendingSnapshot":
        """Returns once every source tensor has been read — with the engine this is after the pack
        kernels, i.e. HBM-speed, not host-link speed; the drain and the commit continue in a thread."""
        cls._validate_app_state(app_state)
        loop = asyncio.new_event_loop()
        pgw = PGWrapper(pg)
        path, globs, keys = cls._coalesce_path_and_replicated(path, pgw, app_state, replicated or [])
        storage = url_to_storage_plugin_in_event_loop(path, loop, storage_options)
        try:
            pending, metadata = cls._take_impl(path, app_state, globs, pgw, storage, loop, True, _custom_tensor_prepare_func, keys)
        except BaseException:
            storage.sync_close(loop)
            loop.close()
            raise
        return PendingSnapshot(path, pending, pgw, metadata, storage, loop, storage_options)

    @classmethod
    def _take_impl(
        cls,
        path: str,
        app_state: AppState,
        replicated: Set[str],
        pgw: PGWrapper,
        storage: StoragePlugin,
        loop: asyncio.AbstractEventLoop,
        is_async_snapshot: bool,
        _custom_tensor_prepare_func: Optional[CustomPrepareFunc],
        global_keys: Optional[List[str]] = None,
    ) -> Tuple[PendingIOWork, SnapshotMetadata]:
        import gc
        import time as _time

        from .scheduler import LAST_STATS

        # the planning below allocates a few thousand small objects; a generational GC pass over a large training heap in
        # the middle of it showed up as 40-50 ms spikes of the async_take blocking window: collect afterwards instead
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            return cls._take_impl_nogc(path, app_state, replicated, pgw, storage, loop, is_async_snapshot, _custom_tensor_prepare_func, LAST_STATS, _time, global_keys)
        finally:
            if gc_was_enabled:
                gc.enable()

    @classmethod
    def _take_impl_nogc(cls, path, app_state, replicated, pgw, storage, loop, is_async_snapshot, _custom_tensor_prepare_func, LAST_STATS, _time, global_keys=None):
        phases: Dict[str, float] = {}
        _t = [_time.perf_counter()]

        def lap(name: str) -> None:
            now = _time.perf_counter()
            phases[name] = phases.get(name, 0.0) + (now - _t[0]) * 1e3
            _t[0] = now

        app_state = dict(app_state)
        rng_item = cls._pop_rng_state(app_state)
        rng_sd = None
        manifest: Manifest = {}
        flattened: Dict[str, Any] = {}
        # capture the RNG state first and put it back after user code ran (T:snapshot.py:538-574)
        if rng_item is not None:
            key, stateful = rng_item
            rng_sd = stateful.state_dict()
            m, f = flatten(rng_sd, prefix=key)
            manifest.update(m)
            flattened.update(f)
        rank = pgw.get_rank()
        # the keys of all ranks normally travel with the path/glob exchange of this take (one collective less)
        if global_keys is None:
            global_keys = cls._gather_keys(list(app_state.keys()), pgw)
        lap("gather_keys")
        for key in global_keys:
            if key in app_state:
                m, f = flatten(app_state[key].state_dict(), prefix=key)
                manifest.update(m)
                flattened.update(f)
            pgw.barrier()  # state_dict() may itself run collectives; keep ranks in step
        if rng_item is not None:
            rng_item[1].load_state_dict(rng_sd)
        lap("state_dict+flatten+barrier")

        replicated_paths = cls._calculate_replicated_entries(flattened, replicated, pgw)
        lap("replicated_paths")
        entries: Dict[str, Entry] = {}
        write_reqs: Dict[str, List[WriteReq]] = {}
        primitives: Dict[str, PrimitiveEntry] = {}
        for logical_path, obj in flattened.items():
            func = functools.partial(_custom_tensor_prepare_func, logical_path) if _custom_tensor_prepare_func is not None else None
            entry, wrs = prepare_write(obj, logical_path, rank, logical_path in replicated_paths, is_async_snapshot, func)
            if isinstance(entry, PrimitiveEntry):
                primitives[logical_path] = entry
            else:
                entries[logical_path] = entry
                write_reqs[logical_path] = wrs
        lap("prepare_write")
        entries, write_reqs = partition_write_reqs(entries, write_reqs, pgw)
        lap("partition")
        flat_reqs = [wr for wrs in write_reqs.values() for wr in wrs]
        if not is_batching_disabled():
            _, flat_reqs = batch_write_requests(list(entries.values()), flat_reqs)
        manifest.update(primitives)
        manifest.update(entries)
        lap("batch")
        manifest = cls._gather_manifest(manifest, pgw)
        lap("gather_manifest")
        budget = get_process_memory_budget_bytes(pgw)
        lap("memory_budget")
        pending = sync_execute_write_reqs(flat_reqs, storage, budget, rank, loop)
        lap("execute_until_staged")
        LAST_STATS["take_phases_ms"] = phases
        metadata = SnapshotMetadata(version=SNAPSHOT_FORMAT_VERSION, world_size=pgw.get_world_size(), manifest=manifest)
        return pending, metadata

    # ---- restore ------------------------------------------------------------------------------------
    def restore(self, app_state: AppState, strict: bool = True) -> None:
        self._validate_app_state(app_state)
        loop = asyncio.new_event_loop()
        pgw = PGWrapper(self.pg)
        storage = url_to_storage_plugin_in_event_loop(self.path, loop, self._storage_options)
        try:
            app_state = dict(app_state)
            rng_item = self._pop_rng_state(app_state)
            for key in self._gather_keys(list(app_state.keys()), pgw):
                self._load_stateful(key, app_state.get(key), strict, storage, pgw, loop)
                pgw.barrier()
            if rng_item is not None:  # last, so nothing run during restore perturbs it
                self._load_stateful(rng_item[0], rng_item[1], strict, storage, pgw, loop)
        finally:
            storage.sync_close(loop)
            loop.close()

    def _load_stateful(
        self, key: str, stateful: Optional[Stateful], strict: bool, storage: StoragePlugin, pgw: PGWrapper, loop: asyncio.AbstractEventLoop
    ) -> None:
        if stateful is None:
            # the other ranks negotiate read-once ranges for this key (one all-gather): take part with nothing to share
            from .scheduler import _plan_shared_reads, _read_once_eligible

            if _read_once_eligible(pgw):
                _plan_shared_reads(pgw, {})
            return
        manifest, merged = get_manifest_for_rank(self.metadata, pgw.get_rank())
        # load straight into the tensors the stateful already owns (no second copy of the state)
        _, flat = flatten(stateful.state_dict(), prefix=key)
        targets = {k: v for k, v in flat.items() if isinstance(v, (torch.Tensor, ShardedTensor, DTensor))}
        handle_sharded_tensor_elasticity(manifest, merged, list(targets.keys()))
        # restore is collective on pgw: replicated ranges may be read once and exchanged GPU to GPU
        state_dict = self._get_state_dict_for_manifest(key, manifest, targets, pgw, storage, loop, shared_pg=pgw)
        if isinstance(stateful, torch.nn.Module):
            stateful.load_state_dict(state_dict, strict=strict)
        else:
            stateful.load_state_dict(state_dict)

    @staticmethod
    def _get_state_dict_for_manifest(
        key: str,
        manifest: Manifest,
        targets: Dict[str, Any],
        pgw: PGWrapper,
        storage: StoragePlugin,
        loop: asyncio.AbstractEventLoop,
        replicate_from_rank0: bool = False,
        shared_pg: Optional[PGWrapper] = None,
    ) -> Any:
        from .flatten import _encode

        root = _encode(key)
        containers: Manifest = {}
        read_reqs: List[ReadReq] = []
        futs = {}
        for logical_path, entry in manifest.items():
            if logical_path.split("/", 1)[0] != root:
                continue
            if is_container_entry(entry):
                containers[logical_path] = entry
                continue
            rrs, fut = prepare_read(entry, targets.pop(logical_path, None))
            read_reqs += rrs
            futs[logical_path] = fut
        if not is_batching_disabled():
            read_reqs = batch_read_requests(read_reqs)
        budget = get_process_memory_budget_bytes(pgw)
        sync_execute_read_reqs(read_reqs, storage, budget, 0 if replicate_from_rank0 else pgw.get_rank(), loop, shared_pg)
        return inflate(containers, {k: f.obj for k, f in futs.items()}, prefix=key)

    def get_state_dict_for_key(self, key: str, replicate_from_rank0: bool = False) -> Any:
        loop = asyncio.new_event_loop()
        pgw = PGWrapper(self.pg)
        rank = 0 if replicate_from_rank0 else pgw.get_rank()
        manifest, _ = get_manifest_for_rank(self.metadata, rank)
        storage = url_to_storage_plugin_in_event_loop(self.path, loop, self._storage_options)
        try:
            return self._get_state_dict_for_manifest(key, manifest, {}, pgw, storage, loop, replicate_from_rank0)
        finally:
            storage.sync_close(loop)
            loop.close()

    def read_object(self, path: str, obj_out: Optional[T] = None, memory_budget_bytes: Optional[int] = None) -> T:
        rank_str, logical = path.split("/", 1)
        manifest, merged = get_manifest_for_rank(self.metadata, int(rank_str))
        if logical not in merged and logical not in manifest:
            raise RuntimeError(
                f'The supplied path "{path}" does not exist in the snapshot\'s manifest. '
                "Please verify the available paths within the snapshot via `snapshot.get_manifest()`."
            )
        entry = merged.get(logical) or manifest[logical]
        if isinstance(entry, PrimitiveEntry):
            return entry.get_value()  # type: ignore[return-value]
        loop = asyncio.new_event_loop()
        pgw = PGWrapper(self.pg)
        storage = url_to_storage_plugin_in_event_loop(self.path, loop, self._storage_options)
        try:
            read_reqs, fut = prepare_read(entry, obj_out, buffer_size_limit_bytes=memory_budget_bytes)
            if not is_batching_disabled():
                read_reqs = batch_read_requests(read_reqs)
            sync_execute_read_reqs(read_reqs, storage, memory_budget_bytes or _MAX_PER_RANK_MEMORY_BUDGET_BYTES, pgw.get_rank(), loop)
        finally:
            storage.sync_close(loop)
            loop.close()
        return fut.obj

    # ---- metadata -----------------------------------------------------------------------------------
    @property
    def metadata(self) -> SnapshotMetadata:
        if self._metadata is None:
            loop = asyncio.new_event_loop()
            storage = url_to_storage_plugin_in_event_loop(self.path, loop, self._storage_options)
            try:
                self._metadata = self._read_snapshot_metadata(storage, loop)
            finally:
                storage.sync_close(loop)
                loop.close()
        return self._metadata

    def get_manifest(self) -> Dict[str, Entry]:
        return copy.deepcopy(self.metadata.manifest)

    @staticmethod
    def _write_snapshot_metadata(
        metadata: SnapshotMetadata, storage: StoragePlugin, loop: asyncio.AbstractEventLoop, encoded: Optional[bytes] = None
    ) -> None:
        buf = encoded if encoded is not None else metadata.to_yaml().encode("utf-8")
        storage.sync_write(WriteIO(path=SNAPSHOT_METADATA_FNAME, buf=buf), loop)

    @staticmethod
    def _read_snapshot_metadata(storage: StoragePlugin, loop: asyncio.AbstractEventLoop) -> SnapshotMetadata:
        rio = ReadIO(path=SNAPSHOT_METADATA_FNAME)
        try:
            storage.sync_read(rio, loop)
        except Exception as e:
            raise RuntimeError(
                f"Failed to read {SNAPSHOT_METADATA_FNAME}. Ensure path to snapshot is correct, "
                "otherwise snapshot is likely incomplete or corrupted."
            ) from e
        return SnapshotMetadata.from_yaml(rio.buf.getvalue().decode("utf-8"))

    # ---- control-plane helpers (object collectives; KB-sized) -----------------------------------------
    @staticmethod
    def _validate_app_state(app_state: AppState) -> None:
        for key, value in app_state.items():
            if not isinstance(value, Stateful):
                raise TypeError(f"Expected Stateful in app_state for key {key}, got {type(value)}.")

    @staticmethod
    def _pop_rng_state(app_state: AppState) -> Optional[Tuple[str, RNGState]]:
        found = [(k, v) for k, v in app_state.items() if isinstance(v, RNGState)]
        if len(found) > 1:
            raise RuntimeError(f"Multiple RNGState objects in app state: {[k for k, _ in found]}")
        if not found:
            return None
        del app_state[found[0][0]]
        return found[0]

    @staticmethod
    def _gather_keys(keys: List[str], pgw: PGWrapper) -> List[str]:
        gathered: List[Any] = [None] * pgw.get_world_size()
        pgw.all_gather_object(gathered, keys)
        return sorted(set(itertools.chain.from_iterable(gathered)))

    @classmethod
    def _coalesce_path_and_replicated(cls, path: str, pgw: PGWrapper, app_state: AppState, replicated: List[str]) -> Tuple[str, Set[str], List[str]]:
        # one exchange carries the path (rank 0's wins, T:snapshot.py:871-877) and the replication globs
        # (T:snapshot.py:880-887); the reference spends a broadcast and an all-gather on them
        globs = cls._infer_replicated(replicated, app_state)
        everyone: List[Any] = [None] * pgw.get_world_size()
        # ... and the app_state keys of every rank (T:snapshot.py:921-927 spends another all-gather on them)
        keys = [k for k, v in app_state.items() if not isinstance(v, RNGState)]  # RNG state is handled apart, like in _take_impl
        pgw.all_gather_object(everyone, (path, globs, keys))
        chosen = everyone[0][0]
        if chosen != path:
            logger.warning(f"Rank {pgw.get_rank()} specified a path ({path}) different from rank 0 ({chosen}). Using path specified by rank 0.")
        return chosen, cls._coalesce_replicated([e[1] for e in everyone]), sorted(set(itertools.chain.from_iterable(e[2] for e in everyone)))

    @staticmethod
    def _coalesce_replicated(global_replicated: List[List[str]]) -> Set[str]:
        """Only globs that every rank specified count (T:snapshot.py:914-918)."""
        return set.intersection(*map(set, global_replicated))

    @staticmethod
    def _infer_replicated(replicated: List[str], app_state: AppState) -> List[str]:
        """DDP-wrapped modules are replicated by construction (T:snapshot.py:897-912)."""
        out = list(replicated)
        if "**" in out:
            return out
        for key, val in app_state.items():
            if isinstance(val, DDP):
                ignored = set(getattr(val, "parameters_to_ignore", []) or [])
                if not ignored:
                    out.append(os.path.join(key, "**"))
                    continue
                for name, _ in itertools.chain(val.named_parameters(), val.named_buffers()):
                    if name not in ignored:
                        out.append(os.path.join(key, name))
        return out

    @staticmethod
    def _calculate_replicated_entries(flattened: Dict[str, Any], replicated: Set[str], pgw: PGWrapper) -> Set[str]:
        mine = [p for p, v in flattened.items() if not is_sharded(v) and any(fnmatch.fnmatch(p, g) for g in replicated)]
        everyone: List[Any] = [None] * pgw.get_world_size()
        pgw.all_gather_object(everyone, mine)
        # replicated only if present on every rank (T:snapshot.py:656-666).  The reference computes this on rank 0 and
        # broadcasts it; the answer is a function of the gathered lists, so each rank derives it locally.
        count: Dict[str, int] = defaultdict(int)
        for paths in everyone:
            for p in set(paths):
                count[p] += 1
        return {p for p, c in count.items() if c == pgw.get_world_size()}

    @staticmethod
    def _gather_manifest(manifest: Dict[str, Entry], pgw: PGWrapper) -> Dict[str, Entry]:
        per_rank: List[Any] = [None] * pgw.get_world_size()
        pgw.all_gather_object(per_rank, manifest)
        per_rank = consolidate_replicated_entries(per_rank)
        return {os.path.join(str(r), p): e for r, m in enumerate(per_rank) for p, e in m.items()}


# ---- async commit ------------------------------------------------------------------------------------
_store_lock = threading.Lock()
# id(pg) -> [pg, store, takes so far].  The entry holds the process group itself, so its id cannot be reused by a
# later group while the entry exists; an entry whose group object differs is stale and replaced.
_store_cache: Dict[int, List[Any]] = {}


def _commit_store(pgw: PGWrapper):
    """(store, tag) for the commit rendezvous of one async_take on `pgw`: a key-value store reachable by all its
    ranks and usable from a background thread (collectives are not), and a tag every rank derives identically —
    the number of async takes issued on THIS process group (a process-global counter would diverge between ranks
    that took a different number of snapshots on other groups).  Rank 0 hosts a TCPStore; its address travels by
    broadcast once per process group."""
    if pgw.get_world_size() == 1:
        return None, ""
    key = id(pgw.pg)
    with _store_lock:
        ent = _store_cache.get(key)
        if ent is not None and ent[0] is pgw.pg:
            ent[2] += 1
            return ent[1], f"tsnap_b200/take{ent[2]}"
    box: List[Any] = [None]
    if pgw.get_rank() == 0:
        sock = socket.socket()
        sock.bind(("", 0))
        port = sock.getsockname()[1]
        sock.close()
        host = os.environ.get("MASTER_ADDR") or socket.gethostname()
        box = [(host, port)]
    pgw.broadcast_object_list(box, src=0)
    host, port = box[0]
    store = dist.TCPStore(host, port, pgw.get_world_size(), pgw.get_rank() == 0, timedelta(seconds=1800), wait_for_workers=False)
    with _store_lock:
        _store_cache[key] = [pgw.pg, store, 0]
    return store, "tsnap_b200/take0"


class PendingSnapshot:
    """Handle returned by :meth:`Snapshot.async_take`; ``wait()`` joins the background drain + commit."""

    DEFAULT_BARRIER_TIMEOUT = timedelta(seconds=1800)

    def __init__(
        self,
        path: str,
        pending_io_work: PendingIOWork,
        pgw: PGWrapper,
        metadata: SnapshotMetadata,
        storage: StoragePlugin,
        loop: asyncio.AbstractEventLoop,
        storage_options: Optional[Dict[str, Any]] = None,
    ) -> None:
        self.path = path
        self.pg = pgw.pg
        self._storage_options = storage_options
        self._metadata = metadata
        self.exc_info: Optional[Any] = None
        self._done = False
        store, tag = _commit_store(pgw)  # collective on first use: must happen on the caller thread
        self.thread = threading.Thread(
            target=self._complete, args=(pending_io_work, pgw.get_rank(), pgw.get_world_size(), metadata, storage, loop, store, tag), daemon=True
        )
        self.thread.start()

    def _complete(self, pending: PendingIOWork, rank: int, world: int, metadata: SnapshotMetadata, storage: StoragePlugin, loop, store, tag: str) -> None:
        # no collectives here: this is not the thread the process group belongs to
        err: Optional[str] = None
        try:
            try:
                pending.sync_complete(loop)
            except Exception as e:
                err = f"rank {rank}: {e}"
                self.exc_info = sys.exc_info()
            if store is None:
                if err is None:
                    Snapshot._write_snapshot_metadata(metadata, storage, loop)
            else:
                # two-phase: everyone reports; rank 0 commits only if all succeeded; everyone learns the verdict
                store.set(f"{tag}/arrive/{rank}", "ok" if err is None else f"err:{err}")
                if rank == 0:
                    keys = [f"{tag}/arrive/{r}" for r in range(world)]
                    store.wait(keys, self.DEFAULT_BARRIER_TIMEOUT)
                    bad = [v for v in (store.get(k).decode() for k in keys) if v != "ok"]
                    if not bad:
                        Snapshot._write_snapshot_metadata(metadata, storage, loop)
                    store.set(f"{tag}/depart", "ok" if not bad else bad[0])
                store.wait([f"{tag}/depart"], self.DEFAULT_BARRIER_TIMEOUT)
                verdict = store.get(f"{tag}/depart").decode()
                # the last rank to learn the verdict clears the rendezvous keys
                if store.add(f"{tag}/left", 1) == world:
                    for k in [f"{tag}/arrive/{r}" for r in range(world)] + [f"{tag}/depart", f"{tag}/left"]:
                        try:
                            store.delete_key(k)
                        except Exception:
                            pass
                if verdict != "ok" and err is None:
                    raise RuntimeError(f"snapshot aborted by a peer: {verdict}")
        except Exception:
            if self.exc_info is None:
                self.exc_info = sys.exc_info()
        finally:
            try:
                storage.sync_close(loop)
                loop.close()
            except Exception:
                pass
            self._done = True

    def wait(self) -> Snapshot:
        self.thread.join()
        if self.exc_info is not None:
            formatted = "".join(traceback.format_exception(*self.exc_info))
            raise RuntimeError(f"Encountered exception while taking snapshot asynchronously:\n{formatted}")
        snap = Snapshot(path=self.path, pg=self.pg, storage_options=self._storage_options)
        snap._metadata = self._metadata
        return snap

    def done(self) -> bool:
        return self._done
