"""The randomised view generator of tests/test_random_views_gpu.py, run through the host executor: same planner
(normalisation, dim merging, granules, tile cover) as the CUDA path, checked against the oracle on the CPU."""
import random

import pytest
import torch

from oracle import ref_port as R
from tests.test_random_views_gpu import random_view
from torchsnapshot_b200 import _native as N


@pytest.mark.parametrize("seed", list(range(8)))
def test_host_pack_and_scatter_random_views(seed):
    rng = random.Random(100 + seed)
    views = [random_view(rng, 5000 * seed + i, "cpu") for i in range(60)]
    off = rng.choice([0, 1, 2, 5, 8, 13])
    total_pad = off
    descs, want = [], []
    for v in views:
        descs.append(N.save_desc(v, off))
        b = R.serialize_view(v)
        want.append((off, b))
        off += len(b)
    info = N.plan_describe(descs, wire_base_align=rng.choice([0, 3, 16, 100]))  # also validates the tile cover
    assert info["bytes_host"] == off - total_pad
    wire = bytearray(off)
    N.host_execute(descs, wire, threads=rng.choice([1, 3]))
    for i, (o, b) in enumerate(want):
        assert bytes(wire[o : o + len(b)]) == b, (seed, i, tuple(views[i].shape), views[i].stride(), views[i].dtype)
    dests, ldescs = [], []
    for v, (o, b) in zip(views, want):
        if 0 in v.stride() and v.numel() > 0:
            dst = torch.zeros(v.shape, dtype=v.dtype)
        else:
            dst = torch.empty_strided(v.shape, v.stride(), dtype=v.dtype)
            dst.zero_()
        dests.append(dst)
        if dst.numel():
            ldescs.append(N.load_desc(dst, o))
    N.host_execute(ldescs, wire, threads=2)
    for i, (dst, (o, b)) in enumerate(zip(dests, want)):
        assert R.serialize_view(dst) == b, (seed, i)


def test_device_classification_of_random_views():
    """The same descriptors, re-labelled as DEVICE copies, must decompose into bulk + LSU tiles that cover every
    byte exactly once (tsnap_plan_describe verifies the cover and fails otherwise)."""
    rng = random.Random(9)
    views = [random_view(rng, 70000 + i, "cpu") for i in range(200)]
    off, descs = 0, []
    for v in views:
        d = N.save_desc(v, off)
        d.src_space = N.SPACE_DEVICE
        d.src_addr = (1 << 34) + (v.data_ptr() & 0xFFFF)  # plausible device address with the same low bits
        descs.append(d)
        off += v.numel() * v.element_size()
    info = N.plan_describe(descs)
    assert info["bytes_bulk"] + info["bytes_lsu"] == off and info["n_members_host"] == 0
