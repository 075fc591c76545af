"""The copy-engine rows kernel (strided members with long 16 B-aligned runs: column shards, narrow on dim != 0,
reshard boxes) and the shared-memory tiled transpose, both directions, against the oracle's serialization."""
import pytest
import torch

from oracle import ref_port as R
from tests.util import det_tensor

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pack(eng, views, offs, total):
    from torchsnapshot_b200 import _native as N

    descs = [N.save_desc(v, o) for v, o in zip(views, offs)]
    sb = eng.stage(descs, total, stream=torch.cuda.current_stream().cuda_stream, keepalive=views)
    got = bytes(sb.wait())
    st = sb.stats()
    sb.release()
    return got, st


def _roundtrip(views, expect_mode):
    """pack the views back to back (16 B-aligned offsets), compare with the oracle, scatter into same-strided twins."""
    from torchsnapshot_b200 import _native as N

    eng = N.get_engine(0)
    offs, off, want = [], 0, []
    for v in views:
        offs.append(off)
        b = R.serialize_view(v)
        want.append(b)
        off += (len(b) + 15) // 16 * 16
    got, st = _pack(eng, views, offs, off)
    for i, (o, b) in enumerate(zip(offs, want)):
        assert got[o : o + len(b)] == b, (i, tuple(views[i].shape), views[i].stride(), views[i].dtype)
    expect_mode(st)
    dests = []
    for v in views:
        d = torch.empty_strided(v.shape, v.stride(), dtype=v.dtype, device=DEV)
        d.zero_()
        dests.append(d)
    eng.consume(got, [N.load_desc(d, o) for d, o in zip(dests, offs)])
    for i, (d, b) in enumerate(zip(dests, want)):
        assert R.serialize_view(d) == b, ("scatter", i)
    return st


def test_rows_kernel_column_shards_and_boxes():
    base = det_tensor((4096, 1024), torch.float32, 1).to(DEV)      # 4 KiB rows
    wide = det_tensor((300, 40000), torch.bfloat16, 2).to(DEV)     # 80 KB rows: runs longer than a 48 KiB stage
    cube = det_tensor((6, 50, 640), torch.float32, 3).to(DEV)
    views = [
        base[:, 128:256],          # 512 B runs in 4 KiB-pitched rows (torchrec COLUMN_WISE shard)
        base[100:3000, 512:],      # 2 KiB runs, row offset
        wide[:, 8:32776],          # 65 536 B runs (> stage size), 16 B-aligned start
        cube[1:5, :, 64:320],      # two outer dims, 1 KiB runs
        base[::2, 256:512],        # stepped rows
    ]

    def rows(st):
        assert st["n_tiles_rows"] > 0 and st["bytes_rows"] == sum(v.numel() * v.element_size() for v in views), st

    _roundtrip(views, rows)


def test_rows_tma_boxes_match_the_per_run_kernel(monkeypatch):
    """Runs of at most 1 KiB move as tensor-map boxes of many runs (clipped at the member's edge); longer runs, 5 outer dims
    or TSNAP_B200_TMA_ROWS=0 stay on the per-run kernel — same bytes either way, both directions."""
    base = det_tensor((5000, 512), torch.float32, 11).to(DEV)          # 2 KiB rows
    deep = det_tensor((3, 2, 3, 2, 37, 96), torch.float32, 12).to(DEV)   # 384 B rows
    views = [
        base[:, 64:128],            # 256 B runs: 128-run boxes, 5000 = 39 x 128 + 8 (a clipped last box)
        base[7:4001, 128:400],      # 1088 B runs: above the crossover, one request per run
        base[:, :],                 # dense: bulk kernel, not a rows member
        base[::3, 0:512:1][:, 256:],  # stepped rows, 1 KiB runs
        deep[::2, :, ::2, 1, 2:35, 16:80],     # 256 B runs under 4 outer dims that do not merge: a rank-5 tensor map
        deep[::2, :, ::2, :, 2:35, 16:80],     # 5 outer dims: more than a tensor map carries, per-run kernel
        deep[1:, :, 1:, :, :, 32:96].transpose(0, 2),
    ]
    stats = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TSNAP_B200_TMA_ROWS", flag)
        stats[flag] = _roundtrip(views, lambda st: None)
    assert stats["1"]["bytes_rows"] == stats["0"]["bytes_rows"] > 0
    assert stats["1"]["n_tiles_rows"] != stats["0"]["n_tiles_rows"], stats  # boxes vs 192 KiB tiles of runs


def test_rows_kernel_declines_short_or_unaligned_runs():
    base = det_tensor((2048, 256), torch.float32, 4).to(DEV)
    views = [base[:, 4:36], base[:, 1:129]]  # 128 B runs (too short); 512 B runs starting 4 B off 16 B alignment

    def lsu(st):
        assert st["n_tiles_rows"] == 0 and st["n_tiles_lsu"] > 0

    _roundtrip(views, lsu)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64, torch.uint8, torch.int64])
def test_tiled_transpose(dtype):
    a = det_tensor((300, 200), dtype, 5).to(DEV)
    b = det_tensor((1000, 1037), dtype, 6).to(DEV)
    c = det_tensor((5, 130, 70), dtype, 7).to(DEV)
    views = [a.t(), b.t(), c.permute(0, 2, 1), c.permute(2, 0, 1), b.t()[3:900, 10:1000], c.transpose(0, 2)]

    def any_mode(st):
        assert st["n_tiles_lsu"] > 0

    _roundtrip(views, any_mode)


def _tma_tiles(sa, sb, esz):
    """tile count of the tensor-map transpose: 32 KiB tiles, the shape variant that pads (sa, sb) least (plan.h)."""
    a0, b0 = (128 if esz == 2 else 64), (64 if esz == 8 else 128)
    return min(-(-sa // a) * -(-sb // b) for a, b in ((a0, b0), (a0 * 2, b0 // 2), (a0 // 2, b0 * 2)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.int32, torch.bfloat16, torch.float16, torch.int16, torch.float64, torch.int64])
def test_tma_transpose_every_tile_shape_and_edges(dtype, monkeypatch):
    """2-D transposes whose strides are 16 B multiples run on the tensor-map TMA kernel: every (element size, tile shape)
    instantiation, clipped edge tiles on both extents, both directions; the LSU transpose is the A/B and the fallback."""
    esz = torch.empty((), dtype=dtype).element_size()
    a0, b0 = (128 if esz == 2 else 64), (64 if esz == 8 else 128)
    lsu_a, lsu_b = (64 if esz >= 4 else 128), (32 if esz == 8 else 64)
    # (rows, cols) of the dense source; the staged view is its .t(): logical (cols, rows), A = dim 0 (extent cols), B = dim 1
    shapes = [(4096, 2048), (4096, a0 // 2), (b0 // 2, 4096), (4096, a0), (b0, 4096), (1000, 1000 + 16 // esz * 3), (520, 264)]
    for seed, (r, c) in enumerate(shapes):
        v = det_tensor((r, c), dtype, 20 + seed).to(DEV).t()
        counts = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("TSNAP_B200_TMA_TRANSPOSE", flag)
            counts[flag] = _roundtrip([v], lambda st: None)["n_tiles_lsu"]
        assert counts["1"] == _tma_tiles(c, r, esz), (dtype, (r, c), counts)
        assert counts["0"] == -(-c // lsu_a) * -(-r // lsu_b), (dtype, (r, c), counts)


def test_tma_transpose_higher_rank_and_fallbacks(monkeypatch):
    """3/4/5-D permutes go through rank-5 tensor maps; 6 dims, 1-byte elements, odd strides or bases stay on the LSU path —
    all byte-identical to the oracle, in one packed image."""
    monkeypatch.setenv("TSNAP_B200_TMA_TRANSPOSE", "1")
    c3 = det_tensor((5, 136, 72), torch.float32, 31).to(DEV)
    c4 = det_tensor((3, 4, 264, 40), torch.bfloat16, 32).to(DEV)
    c5 = det_tensor((2, 3, 2, 72, 24), torch.float64, 33).to(DEV)
    c6 = det_tensor((2, 2, 2, 2, 40, 24), torch.float32, 34).to(DEV)
    odd = det_tensor((257, 1037), torch.float32, 35).to(DEV)
    byt = det_tensor((300, 400), torch.uint8, 36).to(DEV)
    views = [c3.permute(0, 2, 1), c3.permute(2, 0, 1), c3.transpose(0, 2), c4.permute(0, 3, 1, 2), c4.permute(1, 0, 3, 2), c4.transpose(1, 3),
             c5.permute(0, 1, 2, 4, 3), c5.permute(4, 1, 2, 0, 3), c6.permute(0, 1, 2, 3, 5, 4), odd.t(), odd[1:, 4:].t(), byt.t(),
             c3.permute(0, 2, 1)[1:4, 8:64, 16:120]]
    _roundtrip(views, lambda st: None)


def test_transpose_and_rows_through_snapshot_api(tmp_path):
    import torchsnapshot_b200 as B
    from tests.util import wire_bytes

    w = det_tensor((2048, 3072), torch.bfloat16, 8).to(DEV)
    state = {"wt": w.t(), "cols": w[:, 1024:2048], "dense": w}
    snap = B.Snapshot.take(str(tmp_path / "s"), {"m": B.StateDict(**state)})
    tgt = {"wt": torch.zeros(3072, 2048, dtype=torch.bfloat16, device=DEV), "cols": torch.zeros(2048, 3072, dtype=torch.bfloat16, device=DEV)[:, 5:1029],
           "dense": torch.zeros_like(w)}
    sd = B.StateDict(**tgt)
    snap.restore({"m": sd})
    for k in state:
        assert wire_bytes(state[k]) == wire_bytes(sd[k]), k
