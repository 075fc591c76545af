"""Parity of the CUDA path (pack kernels -> pinned ring -> native writes; reads -> H2D -> scatter kernels)
through the public Snapshot API: against the reference's golden fixtures, against the oracle for device
mixes the CPU-only reference run could not produce, and through size-independent properties at
GB scale."""
import json
import os
import tempfile

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu

import torchsnapshot_b200 as B  # noqa: E402
from oracle import ref_port as R  # noqa: E402
from tests.cases import CASES, SHARDED_CASES, apply_knobs, build_sharded  # noqa: E402
from tests.test_parity_cpu import _golden, assert_matches_golden  # noqa: E402
from tests.util import canonicalize, det_tensor, snapshot_digest, wire_bytes  # noqa: E402
from torchsnapshot_b200.flatten import flatten  # noqa: E402

DEV = "cuda:0"
ALL_CUDA = [n for n in sorted(CASES) if n != "model_adam"]


@pytest.fixture(scope="module")
def pg():
    if not dist.is_initialized():
        f = tempfile.NamedTemporaryFile(delete=False)
        dist.init_process_group("gloo", init_method=f"file://{f.name}", rank=0, world_size=1)
    yield
    if dist.is_initialized():
        dist.destroy_process_group()


def _flat_equal(a, b):
    fa, fb = flatten(a, "x")[1], flatten(b, "x")[1]
    assert list(fa) == list(fb)
    for k in fa:
        if isinstance(fa[k], torch.Tensor):
            assert fa[k].device == fb[k].device or True
            assert wire_bytes(fa[k]) == wire_bytes(fb[k]), k
        else:
            assert fa[k] == fb[k], k


@pytest.mark.parametrize("name", ALL_CUDA)
@pytest.mark.parametrize("mode", ["take", "async_take"])
def test_cuda_take_matches_reference_golden(name, mode, tmp_path):
    build, knobs = CASES[name]
    state = build(DEV)
    before = B.get_engine(0).stats()["kernels_launched"]
    with apply_knobs(knobs):
        if mode == "take":
            B.Snapshot.take(str(tmp_path / "snap"), {"state": B.StateDict(**state)})
        else:
            pending = B.Snapshot.async_take(str(tmp_path / "snap"), {"state": B.StateDict(**state)})
            # the sources may be clobbered as soon as async_take returns
            for v in flatten(state, "x")[1].values():
                if isinstance(v, torch.Tensor) and v.numel():
                    v.zero_()
            torch.cuda.synchronize()
            pending.wait()
    assert B.get_engine(0).stats()["kernels_launched"] > before, "the CUDA kernels did not run"
    assert_matches_golden(snapshot_digest(str(tmp_path / "snap")), _golden(name))


def test_mixed_device_state_matches_oracle(tmp_path):
    # Adam keeps `step` on the CPU: CPU and GPU members fill separate slab chains (T:batcher.py:300-303)
    build, knobs = CASES["model_adam"]
    state = build(DEV)
    with apply_knobs(knobs):
        snap = B.Snapshot.take(str(tmp_path / "snap"), {"state": B.StateDict(**state)})
    dig = snapshot_digest(str(tmp_path / "snap"))
    flat = {k: v for k, v in flatten(state, "state")[1].items() if isinstance(v, torch.Tensor)}
    entries, files = R.plan_save(flat, slab_threshold=knobs["slab"])
    manifest, names = canonicalize({f"0/{k}": v for k, v in entries.items()})
    import hashlib

    for path, e in manifest.items():
        assert dig["manifest"][path] == e, path
    for loc, blob in files.items():
        assert dig["files"][names.get(loc, loc)] == {"nbytes": len(blob), "sha256": hashlib.sha256(blob).hexdigest()}, loc
    target = build(DEV)
    for v in flatten(target, "x")[1].values():
        if isinstance(v, torch.Tensor):
            v.zero_()
    tgt = B.StateDict(**target)
    snap.restore({"state": tgt})
    _flat_equal(state, dict(tgt))


@pytest.mark.parametrize("name", ALL_CUDA)
def test_cuda_restore_round_trip_and_cross_device(name, tmp_path):
    build, knobs = CASES[name]
    state = build(DEV)
    with apply_knobs(knobs):
        snap = B.Snapshot.take(str(tmp_path / "snap"), {"state": B.StateDict(**state)})
        for dev in (DEV, "cpu"):  # a GPU-written snapshot restores into GPU and into CPU tensors
            target = build(dev)
            for v in flatten(target, "x")[1].values():
                if isinstance(v, torch.Tensor):
                    v.zero_()
            tgt = B.StateDict(**target)
            snap.restore({"state": tgt})
            _flat_equal(state, dict(tgt))


@pytest.mark.parametrize("name", sorted(SHARDED_CASES))
def test_cuda_sharded_golden_and_reshard(name, tmp_path, pg):
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata

    specs, knobs = SHARDED_CASES[name]
    state = build_sharded(name, DEV)
    with apply_knobs(knobs):
        snap = B.Snapshot.take(str(tmp_path / "snap"), {"state": B.StateDict(**state)})
        assert_matches_golden(snapshot_digest(str(tmp_path / "snap")), _golden(name))
        targets = {}
        for i, (rows, cols, dt, dim, n) in enumerate(specs):
            odim, on = 1 - dim, n + 1
            extent = (rows, cols)[odim]
            step = -(-extent // on)
            shards = []
            for lo in range(0, extent, step):
                ln = min(step, extent - lo)
                off, sz = [0, 0], [rows, cols]
                off[odim], sz[odim] = lo, ln
                shards.append(Shard(tensor=torch.zeros(sz, dtype=dt, device=DEV), metadata=ShardMetadata(shard_offsets=off, shard_sizes=sz, placement=f"rank:0/{DEV}")))
            targets[f"table_{i}"] = ShardedTensor._init_from_local_shards(shards, (rows, cols))
        tgt = B.StateDict(**targets)
        snap.restore({"state": tgt})
    for i, (rows, cols, dt, dim, n) in enumerate(specs):
        full = det_tensor((rows, cols), dt, 900 + i)
        for sh in tgt[f"table_{i}"].local_shards():
            o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
            assert wire_bytes(full[o[0] : o[0] + s[0], o[1] : o[1] + s[1]]) == wire_bytes(sh.tensor)


def test_dtype_mismatch_follows_reference_semantics(tmp_path, pg):
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata

    w = (torch.randn(257, 128, device=DEV) * 5).to(torch.bfloat16)
    snap = B.Snapshot.take(str(tmp_path / "s"), {"state": B.StateDict(w=w)})
    # plain tensors: a dtype mismatch is not an in-place load (T:io_preparers/tensor.py:190-198) — the saved
    # tensor is materialised as saved and handed to load_state_dict
    tgt = B.StateDict(w=torch.zeros(257, 128, device=DEV))
    snap.restore({"state": tgt})
    assert tgt["w"].dtype == torch.bfloat16 and wire_bytes(tgt["w"]) == wire_bytes(w)
    # sharded targets: pieces are copied with Tensor.copy_ semantics, i.e. converted (T:io_preparers/sharded_tensor.py:316-323);
    # here the scatter kernel does the bf16 -> fp32 conversion
    sw = ShardedTensor._init_from_local_shards([Shard(tensor=w.clone(), metadata=ShardMetadata(shard_offsets=[0, 0], shard_sizes=[257, 128], placement=f"rank:0/{DEV}"))], (257, 128))
    snap2 = B.Snapshot.take(str(tmp_path / "s2"), {"state": B.StateDict(w=sw)})
    dst = torch.zeros(257, 128, device=DEV)
    tw = ShardedTensor._init_from_local_shards([Shard(tensor=dst, metadata=ShardMetadata(shard_offsets=[0, 0], shard_sizes=[257, 128], placement=f"rank:0/{DEV}"))], (257, 128))
    snap2.restore({"state": B.StateDict(w=tw)})
    assert torch.equal(dst, w.float())


def test_gb_scale_round_trip_checksums(tmp_path):
    # size-independent property at the scale of one rank of C3 (2 GB): take -> restore is the identity,
    # checked with per-tensor integer checksums computed on the device
    torch.manual_seed(0)
    shapes = [(16032, 4096)] * 2 + [(512, 4096), (128, 4096), (128, 4096), (512, 4096), (1792, 4096), (512, 14336), (1792, 4096), (512,), (512,)] * 16
    state = {f"p{i}": torch.randn(s, device=DEV, dtype=torch.bfloat16) for i, s in enumerate(shapes)}

    def checks(sd):
        return [int(v.view(torch.int16).to(torch.int64).sum().item()) ^ int(v.view(torch.int16)[..., ::7].to(torch.int64).sum().item()) for v in sd.values()]

    want = checks(state)
    snap = B.Snapshot.take(str(tmp_path / "s"), {"state": B.StateDict(**state)})
    total = sum(v.numel() * 2 for v in state.values())
    on_disk = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(tmp_path / "s") for f in fs if f != ".snapshot_metadata")
    assert on_disk == total  # no header, no padding
    tgt = B.StateDict(**{k: torch.zeros_like(v) for k, v in state.items()})
    snap.restore({"state": tgt})
    assert checks(dict(tgt)) == want
    for k in list(state)[:4]:
        assert torch.equal(state[k], tgt[k])


def test_engine_refuses_to_fall_back(monkeypatch):
    # the product path has no CPU/torch fallback for device tensors: a host-only engine rejects them loudly
    from torchsnapshot_b200 import _native as N

    eng = N.Engine(device=-1, io_threads=1, pinned_slot_bytes=1 << 20, pinned_slots=2)
    try:
        t = torch.ones(8, device=DEV)
        with pytest.raises(N.NativeError):
            eng.stage([N.save_desc(t, 0)], 32).wait()
    finally:
        eng.close()


def test_cuda_dtensor_round_trip(tmp_path, pg):
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor, Replicate, Shard

    mesh = init_device_mesh("cuda", (1,))
    w = det_tensor((48, 20), torch.float32, 1).to(DEV)
    e = det_tensor((30, 16), torch.bfloat16, 2).to(DEV)
    state = {
        "w": DTensor.from_local(w.clone(), mesh, [Shard(0)], run_check=False),
        "e": DTensor.from_local(e.clone(), mesh, [Shard(1)], run_check=False),
        "r": DTensor.from_local(det_tensor((7, 5), torch.int64, 3).to(DEV), mesh, [Replicate()], run_check=False),
    }
    with apply_knobs({"max_shard": 1000}):
        snap = B.Snapshot.take(str(tmp_path / "s"), {"m": B.StateDict(**state)})
    ew = snap.get_manifest()["0/m/w"]
    assert [s.sizes for s in ew.shards] == [[12, 20]] * 4
    blob = b"".join(R.serialize_view(w[o : o + 12]) for o in (0, 12, 24, 36))
    (slab,) = os.listdir(tmp_path / "s" / "batched")
    data = (tmp_path / "s" / "batched" / slab).read_bytes()
    lo = ew.shards[0].tensor.byte_range[0]
    assert data[lo : lo + len(blob)] == blob
    tgt = {k: DTensor.from_local(torch.zeros_like(v.to_local()), mesh, v.placements, run_check=False) for k, v in state.items()}
    snap.restore({"m": B.StateDict(**tgt)})
    for k in state:
        assert wire_bytes(state[k].to_local()) == wire_bytes(tgt[k].to_local()), k


def test_read_object_with_memory_budget_on_gpu(tmp_path):
    # budgeted (tiled) reads: T:io_preparers/tensor.py:128-181; strided / offset / prime-sized targets as in the
    # reference's tests/test_tensor_io_preparer.py:159-183
    src = det_tensor((331, 127), torch.float32, 9).to(DEV)
    snap = B.Snapshot.take(str(tmp_path / "s"), {"x": B.StateDict(t=src)})
    for budget in (1000, 4096, 1 << 20):
        out = torch.zeros(331, 127, device=DEV)
        got = snap.read_object("0/x/t", obj_out=out, memory_budget_bytes=budget)
        assert got is out and wire_bytes(out) == wire_bytes(src)
    base = torch.zeros(127, 400, device=DEV)
    strided = base.t()[20:351]  # non-contiguous target with an offset
    got = snap.read_object("0/x/t", obj_out=strided, memory_budget_bytes=5000)
    assert wire_bytes(strided) == wire_bytes(src)


def test_async_take_overlaps_with_compute(tmp_path):
    # C4-style: keep the GPU busy with matmuls while a snapshot drains; the snapshot must hold the values at
    # the time of the call although the parameters are updated right afterwards
    torch.manual_seed(0)
    params = {f"w{i}": torch.randn(2048, 2048, device=DEV) for i in range(24)}  # ~400 MB
    want = {k: v.clone() for k, v in params.items()}
    a = torch.randn(4096, 4096, device=DEV)
    pending = B.Snapshot.async_take(str(tmp_path / "s"), {"m": B.StateDict(**params)})
    for step in range(20):  # "training" continues immediately
        a = (a @ a).clamp_(-1, 1)
        for v in params.values():
            v.add_(1.0)
    torch.cuda.synchronize()
    snap = pending.wait()
    tgt = B.StateDict(**{k: torch.zeros_like(v) for k, v in params.items()})
    snap.restore({"m": tgt})
    for k in want:
        assert torch.equal(want[k], tgt[k]), k


def test_back_to_back_async_takes_do_not_interfere(tmp_path):
    # two snapshots in flight at once: the second must not reuse the staging arena before the first has drained
    a = {f"a{i}": torch.randn(1 << 20, device=DEV) for i in range(48)}  # ~200 MB
    b = {f"b{i}": torch.randn(1 << 20, device=DEV) for i in range(48)}
    want_a = {k: v.clone() for k, v in a.items()}
    want_b = {k: v.clone() for k, v in b.items()}
    p1 = B.Snapshot.async_take(str(tmp_path / "s1"), {"m": B.StateDict(**a)})
    p2 = B.Snapshot.async_take(str(tmp_path / "s2"), {"m": B.StateDict(**b)})
    for v in list(a.values()) + list(b.values()):
        v.zero_()
    s1, s2 = p1.wait(), p2.wait()
    for snap, want in ((s1, want_a), (s2, want_b)):
        tgt = B.StateDict(**{k: torch.empty_like(v) for k, v in want.items()})
        snap.restore({"m": tgt})
        for k in want:
            assert torch.equal(want[k], tgt[k]), k
