"""Randomised parity of the pack / scatter kernels against the oracle: random dtypes, shapes, permutations,
slices with steps, broadcast (stride 0) dims and unpadded slab offsets — every planner mode and alignment
residue shows up.  Seeded, so failures reproduce."""
import random

import pytest
import torch

from oracle import ref_port as R
from tests.util import ALL_RAW_DTYPES, det_tensor

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def random_view(rng: random.Random, seed: int, dev: str) -> torch.Tensor:
    dt = rng.choice(ALL_RAW_DTYPES)
    nd = rng.randint(1, 4)
    shape = [rng.choice([1, 2, 3, 5, 8, 17, 32, 64, 129]) for _ in range(nd)]
    if rng.random() < 0.15:
        shape[rng.randrange(nd)] = rng.choice([1000, 4099, 16384])
    while True:  # keep every base tensor below ~2M elements
        numel = 1
        for s_ in shape:
            numel *= s_
        if numel <= (1 << 21):
            break
        big = max(range(nd), key=lambda d: shape[d])
        shape[big] = max(1, shape[big] // 2)
    t = det_tensor(tuple(shape), dt, seed).to(dev)
    if rng.random() < 0.5 and nd > 1:
        perm = list(range(nd))
        rng.shuffle(perm)
        t = t.permute(perm)
    for d in range(t.dim()):
        if rng.random() < 0.4 and t.shape[d] > 2:
            lo = rng.randrange(0, t.shape[d] - 1)
            hi = rng.randrange(lo + 1, t.shape[d] + 1)
            step = rng.choice([1, 1, 1, 2, 3])
            t = t.narrow(d, lo, hi - lo)
            if step > 1:
                idx = [slice(None)] * t.dim()
                idx[d] = slice(None, None, step)
                t = t[tuple(idx)]
    if rng.random() < 0.1:
        t = t.unsqueeze(0).expand(3, *t.shape)
    if rng.random() < 0.1:
        t = t.reshape(-1)[: max(1, t.numel() // 2)] if t.is_contiguous() else t
    return t


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_slabs_pack_and_scatter(seed):
    from torchsnapshot_b200 import _native as N

    rng = random.Random(seed)
    eng = N.get_engine(0)
    views = [random_view(rng, 1000 * seed + i, DEV) for i in range(90)]
    off, descs, want = rng.choice([0, 0, 1, 3, 8]), [], []
    pad = off
    for v in views:
        descs.append(N.save_desc(v, off))
        b = R.serialize_view(v)
        want.append((off, b))
        off += len(b)
    staged = eng.stage(descs, off, stream=torch.cuda.current_stream().cuda_stream, keepalive=views)
    got = bytes(staged.wait())
    staged.release()
    for i, (o, b) in enumerate(want):
        assert got[o : o + len(b)] == b, f"seed {seed} member {i}: shape={tuple(views[i].shape)} stride={views[i].stride()} dtype={views[i].dtype} off={o}"
    # scatter the same image back into fresh destinations with the views' own striding
    dests, ldescs = [], []
    for v, (o, b) in zip(views, want):
        if 0 in v.stride() and v.numel() > 0:
            dst = torch.zeros(v.shape, dtype=v.dtype, device=DEV)
        else:
            dst = torch.empty_strided(v.shape, v.stride(), dtype=v.dtype, device=DEV)
            dst.zero_()
        dests.append(dst)
        if dst.numel():
            ldescs.append(N.load_desc(dst, o))
    eng.consume(got, ldescs)
    for i, (dst, (o, b)) in enumerate(zip(dests, want)):
        assert R.serialize_view(dst) == b, f"seed {seed} scatter member {i}"


def test_random_reshard_boxes():
    """Random saved-piece x local-shard intersections: the scatter kernel's sub-box copy vs numpy slicing."""
    from torchsnapshot_b200 import _native as N

    rng = random.Random(7)
    eng = N.get_engine(0)
    for case in range(40):
        dt = rng.choice([torch.float32, torch.bfloat16, torch.int64, torch.uint8])
        nd = rng.randint(1, 3)
        gshape = [rng.randint(4, 40) for _ in range(nd)]
        full = det_tensor(tuple(gshape), dt, 50 + case)

        def rand_box():
            off = [rng.randrange(0, s - 1) for s in gshape]
            sz = [rng.randint(1, s - o) for s, o in zip(gshape, off)]
            return off, sz

        s_off, s_sz = rand_box()
        c_off, c_sz = rand_box()
        if not R.boxes_overlap(s_off, s_sz, c_off, c_sz):
            continue
        saved = R.box(full, s_off, s_sz).contiguous()
        buf = R.serialize_view(saved)
        local = torch.zeros(c_sz, dtype=dt, device=DEV)
        region = R.overlap_region(s_off, s_sz, c_off, c_sz)
        dst = local
        first = 0
        strides = [1] * nd
        for i in range(nd - 2, -1, -1):
            strides[i] = strides[i + 1] * s_sz[i + 1]
        for d, so, do, n in region:
            dst = dst.narrow(d, do, n)
            first += so * strides[d]
        eng.consume(buf, [N.load_desc(dst, first * saved.element_size(), wire_strides=strides)])
        ref = torch.zeros(c_sz, dtype=dt)
        R.scatter_bytes(buf, s_sz, dt, [region], ref)
        assert R.serialize_view(local) == R.serialize_view(ref), (case, s_off, s_sz, c_off, c_sz)
