"""Putting the engine underneath an unmodified ``torchsnapshot`` install.

``install()`` swaps the two execution entry points that ``torchsnapshot.snapshot`` calls —
``sync_execute_write_reqs`` (T:snapshot.py:618, T:scheduler.py:342-357) and ``sync_execute_read_reqs``
(T:snapshot.py:810, 485; T:scheduler.py:449-463) — for this package's scheduler.  Planning, manifests, the
partitioner, the batcher's slab assignment, the commit protocol and every storage plugin stay the
reference's; only raw tensor traffic against a local filesystem is rerouted to the engine, everything else
falls through to the stager/consumer objects' own asyncio methods."""
from __future__ import annotations

import importlib
from typing import Any, Dict, Optional

from . import scheduler as _sched

_saved: Dict[str, Any] = {}


def install(torchsnapshot_module: Optional[Any] = None) -> None:
    ts = torchsnapshot_module or importlib.import_module("torchsnapshot")
    snap = importlib.import_module(ts.__name__ + ".snapshot")
    sch = importlib.import_module(ts.__name__ + ".scheduler")
    if _saved:
        return
    _saved.update(
        snap_w=snap.sync_execute_write_reqs,
        snap_r=snap.sync_execute_read_reqs,
        sch_w=sch.sync_execute_write_reqs,
        sch_r=sch.sync_execute_read_reqs,
        sch_ew=sch.execute_write_reqs,
        sch_er=sch.execute_read_reqs,
        modules=(snap, sch),
    )
    snap.sync_execute_write_reqs = _sched.sync_execute_write_reqs
    snap.sync_execute_read_reqs = _sched.sync_execute_read_reqs
    sch.sync_execute_write_reqs = _sched.sync_execute_write_reqs
    sch.sync_execute_read_reqs = _sched.sync_execute_read_reqs
    sch.execute_write_reqs = _sched.execute_write_reqs
    sch.execute_read_reqs = _sched.execute_read_reqs


def uninstall() -> None:
    if not _saved:
        return
    snap, sch = _saved["modules"]
    snap.sync_execute_write_reqs = _saved["snap_w"]
    snap.sync_execute_read_reqs = _saved["snap_r"]
    sch.sync_execute_write_reqs = _saved["sch_w"]
    sch.sync_execute_read_reqs = _saved["sch_r"]
    sch.execute_write_reqs = _saved["sch_ew"]
    sch.execute_read_reqs = _saved["sch_er"]
    _saved.clear()
