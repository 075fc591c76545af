"""``storage_plugins`` entry point (T:storage_plugin.py:56-67): ``b200fs://<path>`` under an UNMODIFIED torchsnapshot.

    pip install torchsnapshot_b200      # registers  [project.entry-points.storage_plugins]  b200fs = ...:b200fs
    torchsnapshot.Snapshot.take("b200fs:///mnt/nvme/ckpt/step100", app_state)

The factory returns a local-filesystem plugin that is a subclass of *torchsnapshot's own* ``StoragePlugin`` ABC (the
reference checks ``isinstance``), serves pickled leaves and ``.snapshot_metadata`` itself, advertises ``native_root``,
and — because choosing this scheme is the opt-in — puts the engine underneath the reference's scheduler
(``torchsnapshot_b200.install()``) so that raw tensor traffic for this root is drained by the pack kernels / pinned
ring / native pwrite workers instead of ``tensor.to('cpu')`` + aiofiles."""
from __future__ import annotations

import importlib
from typing import Any, Dict, Optional

from .fs import FSStoragePlugin

_CLASS_CACHE: Dict[int, type] = {}


def b200fs(path: str, storage_options: Optional[Dict[str, Any]] = None):
    try:
        ref_types = importlib.import_module("torchsnapshot.io_types")
    except ImportError:  # used without the reference installed: this package's own ABC
        return FSStoragePlugin(root=path, storage_options=storage_options)
    base = ref_types.StoragePlugin
    cls = _CLASS_CACHE.get(id(base))
    if cls is None:

        class B200FSStoragePlugin(FSStoragePlugin, base):  # type: ignore[misc, valid-type]
            """torchsnapshot_b200's filesystem plugin, registered with the reference's plugin ABC."""

        cls = _CLASS_CACHE[id(base)] = B200FSStoragePlugin
    from ..integration import install

    install()
    return cls(root=path, storage_options=storage_options)
