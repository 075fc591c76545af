"""GPU parity of the C-ABI data plane against byte-level expectations (pack, D2H, file I/O, H2D, scatter)."""
import os

import pytest
import torch

from tests.util import ALL_RAW_DTYPES, det_tensor, wire_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def N():
    from torchsnapshot_b200 import _native

    return _native


@pytest.fixture(scope="module")
def engine(N):
    eng = N.Engine(device=0, io_threads=8, pinned_slot_bytes=4 << 20, pinned_slots=8)
    yield eng
    eng.close()


def _views(dev):
    out = []
    a = det_tensor((257, 129), torch.float32, 1).to(dev)
    out += [a, a.t(), a[:, 5:77], a[3:200:2], a[::3, ::2], a[2], a[:, 7], a[1:, 1:]]
    b = det_tensor((6, 10, 12, 14), torch.bfloat16, 2).to(dev)
    out += [b, b.permute(3, 1, 0, 2), b[:, 1:4, :, 2:5], b.transpose(1, 2)[1:3], b[..., 1:]]
    c = det_tensor((100003,), torch.uint8, 3).to(dev)
    out += [c, c[1:], c[3:99997], c[::7], c[5:70000]]
    d = det_tensor((1 << 20,), torch.int64, 4).to(dev)
    out += [d, d[1:], d[: (1 << 20) - 3]]
    out += [
        torch.tensor(3.5, device=dev),
        torch.empty(0, device=dev),
        torch.randn(5, 0, 3, device=dev),
        torch.tensor([True, False, True], device=dev),
        torch.randn(3, 1, 5, device=dev).expand(3, 4, 5),
    ]
    for i, dt in enumerate(ALL_RAW_DTYPES):
        out.append(det_tensor((33, 17 + i), dt, 10 + i).to(dev))
        out.append(det_tensor((64, 64), dt, 30 + i).to(dev)[:, 3:50])
    return out


def _pack_expect(views):
    off, descs, exp = 0, [], b""
    return off, descs, exp


@pytest.mark.parametrize("flags", [0, 1])
def test_stage_matches_reference_bytes(N, flags):
    eng = N.Engine(device=0, io_threads=2, pinned_slot_bytes=1 << 20, pinned_slots=4, flags=flags)
    try:
        views = _views("cuda:0")
        off, descs, exp = 0, [], b""
        for v in views:
            descs.append(N.save_desc(v, off))
            e = wire_bytes(v)
            exp += e
            off += len(e)
        sb = eng.stage(descs, off, stream=torch.cuda.current_stream().cuda_stream, keepalive=views)
        got = bytes(sb.wait())
        sb.release()
        assert len(got) == len(exp)
        if got != exp:
            o = 0
            for i, v in enumerate(views):
                n = v.numel() * v.element_size()
                assert got[o : o + n] == exp[o : o + n], f"member {i} shape={tuple(v.shape)} stride={v.stride()} dtype={v.dtype} off={o}"
                o += n
        info = N.plan_describe(descs)
        if flags == 0:
            assert info["n_tiles_bulk"] > 0
    finally:
        eng.close()


def test_save_job_files_and_load_job_roundtrip(N, engine, tmp_path):
    views = _views("cuda:0")
    # several files: slab of all views, plus big single-member files
    big = det_tensor((3000, 4099), torch.float32, 77).to("cuda:0")  # ~49 MB, odd row length
    big_t = big.t()  # strided 49 MB
    job = engine.save_job()
    off, exp = 0, b""
    total = sum(v.numel() * v.element_size() for v in views)
    f_slab = job.add_file(str(tmp_path / "batched" / "slab0"), total)
    for v in views:
        job.add_member(f_slab, N.save_desc(v, off), v)
        e = wire_bytes(v)
        exp += e
        off += len(e)
    f_big = job.add_file(str(tmp_path / "0" / "big"), big.numel() * 4)
    job.add_member(f_big, N.save_desc(big, 0), big)
    f_bigt = job.add_file(str(tmp_path / "0" / "big_t"), big.numel() * 4)
    job.add_member(f_bigt, N.save_desc(big_t, 0), big_t)
    f_empty = job.add_file(str(tmp_path / "0" / "empty"), 0)
    job.submit(torch.cuda.current_stream().cuda_stream)
    job.wait_device()
    job.wait()
    st = job.stats()
    job.destroy()
    assert st["n_kernel_launches"] >= 1
    assert (tmp_path / "batched" / "slab0").read_bytes() == exp
    assert (tmp_path / "0" / "big").read_bytes() == wire_bytes(big)
    assert (tmp_path / "0" / "big_t").read_bytes() == wire_bytes(big_t)
    assert (tmp_path / "0" / "empty").read_bytes() == b""

    # load back into fresh strided destinations
    job = engine.load_job()
    f_slab = job.add_file(str(tmp_path / "batched" / "slab0"), total)
    outs = []
    off = 0
    for v in views:
        if 0 in v.stride() and v.numel() > 0:
            dst = torch.zeros(v.shape, dtype=v.dtype, device=v.device)
        else:
            dst = torch.empty_strided(v.shape, v.stride(), dtype=v.dtype, device=v.device)
            dst.zero_()
        outs.append(dst)
        job.add_member(f_slab, N.load_desc(dst, off), dst)
        off += v.numel() * v.element_size()
    big2 = torch.zeros_like(big)
    f_big = job.add_file(str(tmp_path / "0" / "big"), big.numel() * 4)
    job.add_member(f_big, N.load_desc(big2, 0), big2)
    bigt2 = torch.zeros_like(big).t()
    f_bigt = job.add_file(str(tmp_path / "0" / "big_t"), big.numel() * 4)
    job.add_member(f_bigt, N.load_desc(bigt2, 0), bigt2)
    job.submit(torch.cuda.current_stream().cuda_stream)
    job.wait()
    job.destroy()
    for v, o in zip(views, outs):
        assert wire_bytes(v) == wire_bytes(o), (v.shape, v.stride(), v.dtype)
    assert wire_bytes(big) == wire_bytes(big2)
    assert wire_bytes(big_t) == wire_bytes(bigt2)


def test_reshard_box_scatter(N, engine, tmp_path):
    # saved piece [40, 64] fp32; destination receives the sub-box rows 10:30, cols 8:40 into a
    # narrowed view of a bigger local shard (reshard-on-load overlap, sharded_tensor.py:285-298)
    piece = det_tensor((40, 64), torch.float32, 5).to("cuda:0")
    path = str(tmp_path / "piece")
    job = engine.save_job()
    f = job.add_file(path, piece.numel() * 4)
    job.add_member(f, N.save_desc(piece, 0), piece)
    job.submit(torch.cuda.current_stream().cuda_stream)
    job.wait()
    job.destroy()
    local = torch.zeros(50, 100, device="cuda:0")
    dst = local[5:25, 20:52]
    job = engine.load_job()
    f = job.add_file(path, piece.numel() * 4)
    wire_off = (10 * 64 + 8) * 4
    job.add_member(f, N.load_desc(dst, wire_off, wire_strides=[64, 1]), local)
    job.submit()
    job.wait()
    job.destroy()
    ref = torch.zeros(50, 100)
    ref[5:25, 20:52] = piece.cpu()[10:30, 8:40]
    assert wire_bytes(local) == wire_bytes(ref)


def test_multi_wave_small_arena(N, tmp_path):
    eng = N.Engine(device=0, io_threads=4, pinned_slot_bytes=1 << 20, pinned_slots=3, hbm_staging_bytes=8 << 20)
    try:
        tensors = [det_tensor((700_000 + 1000 * i,), torch.float32, 100 + i).to("cuda:0") for i in range(9)]  # ~2.8 MB each
        job = eng.save_job()
        for i, t in enumerate(tensors):
            f = job.add_file(str(tmp_path / f"t{i}"), t.numel() * 4)
            job.add_member(f, N.save_desc(t, 0), t)
        job.submit(torch.cuda.current_stream().cuda_stream)
        job.wait()
        job.destroy()
        for i, t in enumerate(tensors):
            assert (tmp_path / f"t{i}").read_bytes() == wire_bytes(t)
        outs = [torch.zeros_like(t) for t in tensors]
        job = eng.load_job()
        for i, t in enumerate(outs):
            f = job.add_file(str(tmp_path / f"t{i}"), t.numel() * 4)
            job.add_member(f, N.load_desc(t, 0), t)
        job.submit()
        job.wait()
        job.destroy()
        for a, b in zip(tensors, outs):
            assert wire_bytes(a) == wire_bytes(b)
        assert eng.stats()["hbm_arena_bytes"] <= 8 << 20
    finally:
        eng.close()


@pytest.mark.parametrize("src,dst", [(torch.float32, torch.bfloat16), (torch.float32, torch.float16), (torch.bfloat16, torch.float32), (torch.float64, torch.float32), (torch.float16, torch.float64)])
def test_fused_cast_matches_torch(N, engine, src, dst):
    x = (torch.randn(513, 257, dtype=torch.float64) * 3).to(src).to("cuda:0")
    for v in (x, x[:, 3:200], x.t()):
        n = v.numel() * torch.empty(0, dtype=dst).element_size()
        sb = engine.stage([N.save_desc(v, 0, wire_dtype=dst)], n, stream=torch.cuda.current_stream().cuda_stream, keepalive=[v])
        got = bytes(sb.wait())
        sb.release()
        assert got == wire_bytes(v.to(dst)), (src, dst, v.stride())


def test_consume_scatter(N, engine):
    t = det_tensor((123, 77), torch.bfloat16, 9)
    buf = wire_bytes(t)
    dst = torch.zeros(77, 123, dtype=torch.bfloat16, device="cuda:0").t()
    engine.consume(buf, [N.load_desc(dst, 0)])
    assert wire_bytes(dst) == buf


def test_large_throughput_smoke(N, engine, tmp_path):
    # 1 GiB contiguous + odd offsets; checks a checksum rather than bytes, prints kernel timing
    t = torch.empty(1 << 28, dtype=torch.float32, device="cuda:0").uniform_()
    s = torch.ones(3, dtype=torch.uint8, device="cuda:0")
    job = engine.save_job()
    f = job.add_file(str(tmp_path / "slab"), 3 + t.numel() * 4)
    job.add_member(f, N.save_desc(s, 0), s)
    job.add_member(f, N.save_desc(t, 3), t)  # destination misaligned by 3 bytes
    f2 = job.add_file(str(tmp_path / "aligned"), t.numel() * 4)
    job.add_member(f2, N.save_desc(t, 0), t)
    job.submit(torch.cuda.current_stream().cuda_stream)
    job.wait()
    st = job.stats()
    job.destroy()
    print("save stats", st)
    out = torch.zeros_like(t)
    job = engine.load_job()
    f = job.add_file(str(tmp_path / "slab"), t.numel() * 4, offset=3)
    job.add_member(f, N.load_desc(out, 0), out)
    job.submit()
    job.wait()
    print("load stats", job.stats())
    job.destroy()
    assert torch.equal(out, t)
    out.zero_()
    job = engine.load_job()
    f = job.add_file(str(tmp_path / "aligned"), t.numel() * 4)
    job.add_member(f, N.load_desc(out, 0), out)
    job.submit()
    job.wait()
    job.destroy()
    assert torch.equal(out, t)


def test_trim_releases_hbm_and_pinned_memory(N, tmp_path, monkeypatch):
    # the engine's OWN arena (what a C caller gets; Python jobs normally lend one from PyTorch's allocator)
    monkeypatch.setenv("TSNAP_B200_ENGINE_ARENA", "1")
    monkeypatch.setenv("TSNAP_B200_KEEP_ARENA", "1")
    eng = N.Engine(device=0, io_threads=4, pinned_slot_bytes=4 << 20, pinned_slots=4)
    try:
        t = det_tensor((1 << 22,), torch.float32, 3).to("cuda:0")  # 16 MiB
        for rep in range(2):
            job = eng.save_job()
            f = job.add_file(str(tmp_path / f"t{rep}"), t.numel() * 4)
            job.add_member(f, N.save_desc(t, 0), t)
            job.submit(torch.cuda.current_stream().cuda_stream)
            job.wait()
            job.destroy()
            assert eng.stats()["hbm_arena_bytes"] >= t.numel() * 4
            eng.trim()
            assert eng.stats()["hbm_arena_bytes"] == 0
            assert (tmp_path / f"t{rep}").read_bytes() == wire_bytes(t)
    finally:
        eng.close()


def test_column_shard_throughput_smoke(N, engine):
    # column-wise shard (narrow on dim 1): rows of 512 B inside 1 KiB-pitched storage -> STRIDED mode, 16 B granules
    base = torch.empty(1 << 20, 256, dtype=torch.float32, device="cuda:0").uniform_()  # 1 GiB
    view = base[:, 64:192]  # 512 MiB payload
    n = view.numel() * 4
    for _ in range(2):
        sb = engine.stage([N.save_desc(view, 0)], n, stream=torch.cuda.current_stream().cuda_stream, keepalive=[base])
        mv = sb.wait()
        st = sb.stats()
        got = torch.frombuffer(mv, dtype=torch.float32).reshape(view.shape).clone()
        del mv
        sb.release()
    ms = st["kernel_lsu_ms"] + st["kernel_rows_ms"]
    print("column shard stats", {k: st[k] for k in ("kernel_lsu_ms", "kernel_rows_ms", "n_tiles_lsu", "n_tiles_rows")}, "GB/s", 2 * n / 1e9 / (ms / 1e3))
    assert st["n_tiles_rows"] > 0, "512 B runs at 16 B alignment belong to the copy-engine rows kernel"
    assert torch.equal(got, view.cpu())
