"""(The reference's own async_take cannot run in this container: T:dist_store.py:69 unpacks a 4-tuple from an
IPv4 getsockname(); SURVEY.md §4.  async_take is covered through this package's Snapshot instead.)

`torchsnapshot_b200.install()` puts the engine underneath the UNMODIFIED reference: its own
Snapshot.take/restore keep producing the golden bytes, but the raw tensor traffic goes through
libtsnap_b200.so.  Needs the reference tree (build container only)."""
import os
import sys

import pytest
import torch

from tests.cases import CASES, apply_knobs
from tests.test_parity_cpu import _golden, assert_matches_golden
from tests.util import snapshot_digest, wire_bytes

HAVE_REF = os.path.isdir("/root/reference/torchsnapshot")
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")


@pytest.fixture()
def ref():
    sys.path.insert(0, "/root/reference")
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    try:
        import torchsnapshot
    finally:
        sys.path.remove("/root/reference")
    import torchsnapshot_b200 as B

    B.install(torchsnapshot)
    yield torchsnapshot
    B.uninstall()


@pytest.mark.parametrize("name", ["dtypes_and_views", "chunked_slabbed", "slabs", "model_adam"])
def test_reference_api_runs_on_the_engine(ref, name, tmp_path):
    import torchsnapshot_b200 as B
    from torchsnapshot_b200.flatten import flatten

    build, knobs = CASES[name]
    state = build("cpu")
    eng = B.get_engine(-1)
    before = eng.stats()["bytes_written"]
    with apply_knobs(knobs):
        snap = ref.Snapshot.take(str(tmp_path / "snap"), {"state": ref.StateDict(**state)})
        assert eng.stats()["bytes_written"] > before, "the engine was not on the path"
        assert_matches_golden(snapshot_digest(str(tmp_path / "snap")), _golden(name))
        target = build("cpu")
        for v in flatten(target, "x")[1].values():
            if isinstance(v, torch.Tensor):
                v.zero_()
        before_r = eng.stats()["bytes_read"]
        tgt = ref.StateDict(**target)
        snap.restore({"state": tgt})
        assert eng.stats()["bytes_read"] > before_r
    a, b = flatten(state, "x")[1], flatten(dict(tgt), "x")[1]
    for k in a:
        if isinstance(a[k], torch.Tensor):
            assert wire_bytes(a[k]) == wire_bytes(b[k]), k
        else:
            assert a[k] == b[k], k
