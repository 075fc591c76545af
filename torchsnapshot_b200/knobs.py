"""Tunables of the save/restore path.  Same environment variables as the reference (T:knobs.py:23-38,
T:scheduler.py:48) so existing deployments keep their overrides; plus the engine's own TSNAP_B200_* knobs
(read in _native.get_engine)."""
from __future__ import annotations

import os
from contextlib import contextmanager
from typing import Iterator, Optional

MiB = 1024 * 1024


class _IntKnob:
    def __init__(self, env: str, default: int) -> None:
        self.env = env
        self.default = default

    def get(self) -> int:
        raw = os.environ.get(self.env)
        return self.default if raw is None else int(raw)

    @contextmanager
    def override(self, value: int) -> Iterator[None]:
        old = os.environ.get(self.env)
        os.environ[self.env] = str(value)
        try:
            yield
        finally:
            if old is None:
                os.environ.pop(self.env, None)
            else:
                os.environ[self.env] = old


class _FlagKnob(_IntKnob):
    def get(self) -> bool:  # type: ignore[override]
        return os.environ.get(self.env, "False").lower() in ("true", "1")


# tensors above this are written as several chunk files (T:io_preparer.py:122, T:chunked_tensor.py:43)
MAX_CHUNK_SIZE = _IntKnob("TORCHSNAPSHOT_MAX_CHUNK_SIZE_BYTES_OVERRIDE", 512 * MiB)
# local shards above this are subdivided along the sharding dim (T:sharded_tensor.py:48-78)
MAX_SHARD_SIZE = _IntKnob("TORCHSNAPSHOT_MAX_SHARD_SIZE_BYTES_OVERRIDE", 512 * MiB)
# first-fit slab threshold of the batcher (T:batcher.py:246-248)
SLAB_SIZE_THRESHOLD = _IntKnob("TORCHSNAPSHOT_SLAB_SIZE_THRESHOLD_BYTES_OVERRIDE", 128 * MiB)
MAX_IO_CONCURRENCY = _IntKnob("TORCHSNAPSHOT_MAX_PER_RANK_IO_CONCURRENCY_OVERRIDE", 16)
DISABLE_BATCHING = _FlagKnob("TORCHSNAPSHOT_DISABLE_BATCHING", 0)
ELASTICITY_ROOT_ONLY = _FlagKnob("TORCHSNAPSHOT_ENABLE_SHARDED_TENSOR_ELASTICITY_ROOT_ONLY", 0)
MEMORY_BUDGET_ENV = "TORCHSNAPSHOT_PER_RANK_MEMORY_BUDGET_BYTES"


def get_max_chunk_size_bytes() -> int:
    return MAX_CHUNK_SIZE.get()


def get_max_shard_size_bytes() -> int:
    return MAX_SHARD_SIZE.get()


def get_slab_size_threshold_bytes() -> int:
    return SLAB_SIZE_THRESHOLD.get()


def get_max_per_rank_io_concurrency() -> int:
    return MAX_IO_CONCURRENCY.get()


def is_batching_disabled() -> bool:
    return DISABLE_BATCHING.get()


def is_sharded_tensor_elasticity_enabled_at_root_only() -> bool:
    return ELASTICITY_ROOT_ONLY.get()


def get_memory_budget_override() -> Optional[int]:
    raw = os.environ.get(MEMORY_BUDGET_ENV)
    if raw is None:
        return None
    try:
        return int(raw)
    except ValueError:
        return None


override_max_chunk_size_bytes = MAX_CHUNK_SIZE.override
override_max_shard_size_bytes = MAX_SHARD_SIZE.override
override_slab_size_threshold_bytes = SLAB_SIZE_THRESHOLD.override
override_max_per_rank_io_concurrency = MAX_IO_CONCURRENCY.override


@contextmanager
def override_is_batching_disabled(disabled: bool) -> Iterator[None]:
    with DISABLE_BATCHING.override(int(bool(disabled))):
        yield
