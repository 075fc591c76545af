import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_native_library() -> None:
    """The test-suite needs torchsnapshot_b200/lib/libtsnap_b200.so; build it (nvcc cross-compiles without a GPU) when
    a fresh checkout has not run __graft_entry__.build() yet."""
    lib = os.path.join(ROOT, "torchsnapshot_b200", "lib", "libtsnap_b200.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.run(["make", "-C", os.path.join(ROOT, "torchsnapshot_b200", "csrc")], check=True)


def pytest_configure(config):
    _ensure_native_library()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

# tests/ref_suite holds the runner for the reference's own test files; it is driven by tests/test_reference_suite.py
collect_ignore = ["ref_suite"]
