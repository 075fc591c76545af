"""The C-ABI library loads on a CPU-only box and exports every symbol include/tsnap_b200.h declares."""
import ctypes
import os
import re

from torchsnapshot_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    header = open(os.path.join(ROOT, "include", "tsnap_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*|size_t)\s+(tsnap_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_native.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_dtype_sizes():
    assert _native.lib.tsnap_abi_version() == 2
    import torch

    for dt, code in _native.TORCH_TO_TSNAP.items():
        assert _native.lib.tsnap_dtype_size(code) == torch.empty(0, dtype=dt).element_size()


def test_struct_layouts_match_header():
    # sizes the C side assumes (checked indirectly: a wrong layout breaks every parity test, this localises it)
    assert ctypes.sizeof(_native.CopyDesc) == 8 + 8 + 3 * 8 * 8 + 6 * 4 + 8 + 8
    assert ctypes.sizeof(_native.EngineConfig) == 32
    assert ctypes.sizeof(_native.TraceRec) == 40
    assert ctypes.sizeof(_native.ArenaHint) == 32
    assert ctypes.sizeof(_native.JobStats) == 28 * 8


def test_device_engine_fails_loudly_without_gpu():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.NativeError, match="not usable"):
        _native.Engine(device=0)


def test_planner_tile_cover_and_classification():
    import torch

    N = _native
    # describe device-space copies without touching a device: addresses are only classified
    def dev_desc(numel, esz_dtype, src_addr, dst_off):
        d = N.CopyDesc()
        d.ndim = 1
        d.sizes[0] = numel
        d.src_strides[0] = 1
        d.src_addr = src_addr
        d.dst_addr = dst_off
        d.src_dtype = d.dst_dtype = esz_dtype
        d.src_space, d.dst_space = N.SPACE_DEVICE, N.SPACE_WIRE
        return d

    base = 1 << 33
    info = N.plan_describe([dev_desc(1 << 20, N.F32, base, 0)])  # 4 MiB dense aligned -> bulk
    assert info["n_members_bulk"] == 1 and info["n_tiles_bulk"] == 22 and info["n_tiles_lsu"] == 0
    info = N.plan_describe([dev_desc((1 << 20) + 3, N.U8, base, 0)])  # bulk body + 3-byte tail
    assert info["n_members_bulk"] == 1 and info["n_members_lsu"] == 1 and info["bytes_lsu"] == 3
    info = N.plan_describe([dev_desc(1 << 20, N.F32, base, 3)])  # destination misaligned -> LSU contig
    assert info["n_members_bulk"] == 0 and info["n_tiles_lsu"] == 33
    info = N.plan_describe([dev_desc(100, N.F32, base, 0)])  # too small for the bulk engine
    assert info["n_members_bulk"] == 0 and info["n_tiles_lsu"] == 1
    # strided 2-D view
    d = N.CopyDesc()
    d.ndim = 2
    d.sizes[0], d.sizes[1] = 1000, 64
    d.src_strides[0], d.src_strides[1] = 256, 1
    d.src_addr, d.dst_addr = base, 0
    d.src_dtype = d.dst_dtype = N.BF16
    d.src_space, d.dst_space = N.SPACE_DEVICE, N.SPACE_WIRE
    info = N.plan_describe([d])
    assert info["n_members_lsu"] == 1 and info["bytes_lsu"] == 1000 * 64 * 2
