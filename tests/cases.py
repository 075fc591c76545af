"""Deterministic app-state cases shared by the golden generator (run against the unmodified reference,
``oracle/gen_golden.py``) and the parity tests (run against this package, on CPU and on the GPU).

A case is ``name -> (build(device) -> nested dict of leaves, knobs)``; knobs are environment overrides
that shrink the reference's size thresholds so that KB-sized tensors exercise the GB-sized code paths
(the reference's own tests do the same, e.g. tests/test_ddp.py:37-46)."""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Callable, Dict, Tuple

import torch

from tests.util import ALL_RAW_DTYPES, det_tensor

KNOB_ENV = {
    "max_chunk": "TORCHSNAPSHOT_MAX_CHUNK_SIZE_BYTES_OVERRIDE",
    "max_shard": "TORCHSNAPSHOT_MAX_SHARD_SIZE_BYTES_OVERRIDE",
    "slab": "TORCHSNAPSHOT_SLAB_SIZE_THRESHOLD_BYTES_OVERRIDE",
    "no_batching": "TORCHSNAPSHOT_DISABLE_BATCHING",
}


def _dtypes_and_views(dev: str) -> Dict[str, Any]:
    out: Dict[str, Any] = OrderedDict()
    for i, dt in enumerate(ALL_RAW_DTYPES):
        name = str(dt).split(".")[1]
        base = det_tensor((12, 18), dt, 100 + i).to(dev)
        out[f"{name}_dense"] = base
        out[f"{name}_t"] = base.t()
        out[f"{name}_cols"] = base[:, 3:11]
        out[f"{name}_rows"] = base[2:9]
        out[f"{name}_step"] = base[::2, 1::3]
        if dt != torch.bfloat16:
            # the reference truncates / rejects CPU bfloat16 tensors with an odd element count (its
            # untyped-storage trick re-types the bytes as float32, T:serialization.py:208-230), so those
            # shapes cannot be pinned by a golden; they are covered against the oracle instead
            out[f"{name}_scalar"] = det_tensor((), dt, 200 + i).to(dev)
            out[f"{name}_one"] = det_tensor((1,), dt, 300 + i).to(dev)
            out[f"{name}_odd"] = det_tensor((7,), dt, 400 + i).to(dev)[1:6]
        else:
            out[f"{name}_odd"] = det_tensor((8,), dt, 400 + i).to(dev)[1:7]
    out["prims"] = {"i": 7, "f": 0.25, "s": "hello/world%", "b": True, "raw": b"\x00\x01\xfe"}
    out["nested"] = [det_tensor((3, 3), torch.float32, 1).to(dev), {"k/1": det_tensor((5,), torch.int64, 2).to(dev), 3: 4}]
    return out


def _chunked(dev: str) -> Dict[str, Any]:
    out: Dict[str, Any] = OrderedDict()
    out["small"] = det_tensor((7, 10), torch.float32, 1).to(dev)  # 280 B: not chunked at 1000
    out["rows"] = det_tensor((100, 30), torch.float32, 2).to(dev)  # 12000 B -> 12 chunks
    out["ragged"] = det_tensor((37, 11), torch.float64, 3).to(dev)  # 3256 B -> 4 requested chunks of 10 rows
    out["noncontig"] = det_tensor((40, 40), torch.float64, 4).to(dev).t()
    out["cube_t"] = det_tensor((16, 12, 20), torch.bfloat16, 5).to(dev).permute(2, 0, 1)
    out["vec"] = det_tensor((3001,), torch.uint8, 6).to(dev)
    out["wide_row"] = det_tensor((3, 2000), torch.int16, 7).to(dev)  # one row already exceeds the limit
    return out


def _slabs(dev: str) -> Dict[str, Any]:
    out: Dict[str, Any] = OrderedDict()
    sizes = [1000, 3000, 96, 4000, 4096, 5000, 1, 2047, 2048, 1, 4095, 16, 512, 3584, 8, 8192, 100, 100, 3896, 7]
    for i, n in enumerate(sizes):
        out[f"u8_{i}"] = det_tensor((n,), torch.uint8, 50 + i).to(dev)
    for i in range(12):
        out[f"mix_{i}"] = det_tensor((17 + i, 10), [torch.float32, torch.bfloat16, torch.int64][i % 3], 90 + i).to(dev)
    return out


def _model_adam(dev: str) -> Dict[str, Any]:
    # hand-rolled "state dicts" with the structure torch.nn / torch.optim produce
    params = OrderedDict()
    shapes = [("conv.weight", (16, 3, 3, 3)), ("conv.bias", (16,)), ("bn.weight", (16,)), ("bn.bias", (16,)),
              ("bn.running_mean", (16,)), ("bn.running_var", (16,)), ("fc.weight", (10, 144)), ("fc.bias", (10,))]
    for i, (n, s) in enumerate(shapes):
        params[n] = det_tensor(s, torch.float32, 500 + i).to(dev)
    params["bn.num_batches_tracked"] = torch.tensor(3, dtype=torch.int64, device=dev)
    state = {}
    for i, (n, s) in enumerate(shapes[:4] + shapes[6:]):
        state[i] = {
            "step": torch.tensor(float(i + 1)),  # Adam keeps `step` on the CPU by default
            "exp_avg": det_tensor(s, torch.float32, 600 + i).to(dev),
            "exp_avg_sq": det_tensor(s, torch.float32, 700 + i).to(dev),
        }
    groups = [{"lr": 0.001, "betas": (0.9, 0.999), "eps": 1e-08, "weight_decay": 0, "amsgrad": False, "params": list(range(6))}]
    return OrderedDict(model=params, optim={"state": state, "param_groups": groups})


CASES: Dict[str, Tuple[Callable[[str], Dict[str, Any]], Dict[str, int]]] = {
    "dtypes_and_views": (_dtypes_and_views, {}),
    "dtypes_no_batching": (_dtypes_and_views, {"no_batching": 1}),
    "chunked": (_chunked, {"max_chunk": 1000}),
    "chunked_slabbed": (_chunked, {"max_chunk": 1000, "slab": 2048}),
    "slabs": (_slabs, {"slab": 4096}),
    "model_adam": (_model_adam, {"slab": 2048}),
}


# ShardedTensor cases: [(global shape, dtype, sharding dim, n shards)], all shards placed on rank 0
SHARDED_CASES = {
    "sharded_dim0": ([(64, 24, torch.float32, 0, 4), (50, 8, torch.bfloat16, 0, 3)], {"max_shard": 700}),
    "sharded_dim1": ([(20, 96, torch.float32, 1, 4), (9, 35, torch.int64, 1, 5)], {"max_shard": 600}),
}


def build_sharded(case: str, dev: str, init_from_seed: bool = True):
    """name -> ShardedTensor built from local shards living on `dev` (needs an initialised process group)."""
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata

    specs, _ = SHARDED_CASES[case]
    out = OrderedDict()
    for i, (rows, cols, dt, dim, n) in enumerate(specs):
        full = det_tensor((rows, cols), dt, 900 + i) if init_from_seed else torch.zeros((rows, cols), dtype=dt)
        extent = (rows, cols)[dim]
        step = -(-extent // n)
        shards = []
        for lo in range(0, extent, step):
            ln = min(step, extent - lo)
            off = [0, 0]
            off[dim] = lo
            sz = [rows, cols]
            sz[dim] = ln
            local = full.narrow(dim, lo, ln).contiguous().to(dev)
            shards.append(Shard(tensor=local, metadata=ShardMetadata(shard_offsets=off, shard_sizes=sz, placement=f"rank:0/{dev}")))
        out[f"table_{i}"] = ShardedTensor._init_from_local_shards(shards, (rows, cols))
    return out


def apply_knobs(knobs: Dict[str, int]):
    """Context manager setting the env overrides of a case."""
    import contextlib
    import os

    @contextlib.contextmanager
    def cm():
        old = {}
        for k, v in knobs.items():
            env = KNOB_ENV[k]
            old[env] = os.environ.get(env)
            os.environ[env] = str(v)
        try:
            yield
        finally:
            for env, v in old.items():
                if v is None:
                    os.environ.pop(env, None)
                else:
                    os.environ[env] = v

    return cm()
