"""URL -> StoragePlugin resolution (T:storage_plugin.py:20-80).  ``fs://`` / bare paths resolve to the
native-aware FSStoragePlugin; any other scheme is looked up in the ``storage_plugins`` entry-point
group exactly like the reference, so third-party plugins keep working (through the stager seam).
The reference's s3/gcs plugins are cloud I/O and out of scope here."""
from __future__ import annotations

import asyncio
from typing import Any, Dict, Optional

from .io_types import StoragePlugin
from .storage_plugins.fs import FSStoragePlugin


def url_to_storage_plugin(url_path: str, storage_options: Optional[Dict[str, Any]] = None) -> StoragePlugin:
    if "://" in url_path:
        protocol, path = url_path.split("://", 1)
        if len(protocol) == 0:
            protocol = "fs"
    else:
        protocol, path = "fs", url_path
    if storage_options is None:
        storage_options = {}
    if protocol == "fs":
        return FSStoragePlugin(root=path, storage_options=storage_options)
    from importlib.metadata import entry_points

    for ep in entry_points(group="storage_plugins"):
        if ep.name == protocol:
            plugin = ep.load()(path, storage_options)
            if not isinstance(plugin, StoragePlugin):
                raise RuntimeError(f"The factory function for {protocol} ({ep.value}) returned {type(plugin)}.")
            return plugin
    raise RuntimeError(f"Unsupported protocol: {protocol}.")


def url_to_storage_plugin_in_event_loop(
    url_path: str, event_loop: asyncio.AbstractEventLoop, storage_options: Optional[Dict[str, Any]] = None
) -> StoragePlugin:
    async def make() -> StoragePlugin:
        return url_to_storage_plugin(url_path, storage_options)

    return event_loop.run_until_complete(make())
