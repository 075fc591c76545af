"""torchsnapshot_b200 — a B200-native data plane behind TorchSnapshot's Snapshot.take / async_take /
restore and StoragePlugin API.  The device->host drain + serialization (save) and its mirror (restore)
run as hand-written sm_100a kernels plus a native pinned-memory copy/I-O engine (libtsnap_b200.so);
the on-disk format is the reference's, byte for byte."""
from ._native import NativeError, get_engine
from .integration import install, uninstall
from .prepare import cast_on_save, quantize_on_save
from .snapshot import PendingSnapshot, Snapshot
from .stateful import AppState, RNGState, StateDict, Stateful

__version__ = "0.1.0"
__all__ = ["Snapshot", "PendingSnapshot", "Stateful", "StateDict", "RNGState", "AppState", "NativeError", "get_engine", "install", "uninstall", "cast_on_save", "quantize_on_save", "__version__"]
