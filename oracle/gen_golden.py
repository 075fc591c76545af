"""Generates tests/golden/*.json by running the UNMODIFIED reference (imported from /root/reference)
on the deterministic cases of tests/cases.py.  Run here (CPU container):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

The reference cannot travel to the GPU box; its outputs do, as these fixtures: canonical manifest +
sha256 of every raw payload file.  Payload bytes do not depend on the device the tensors live on, so the
same fixtures pin the CUDA path."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torchsnapshot  # noqa: E402  (the reference)

from tests.cases import CASES, SHARDED_CASES, apply_knobs, build_sharded  # noqa: E402
from tests.util import snapshot_digest  # noqa: E402

assert torchsnapshot.__file__.startswith("/root/reference"), torchsnapshot.__file__
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def run(name, app_state, knobs):
    with tempfile.TemporaryDirectory() as d, apply_knobs(knobs):
        path = os.path.join(d, "snap")
        torchsnapshot.Snapshot.take(path, app_state)
        dig = snapshot_digest(path)
    dig["knobs"] = knobs
    dig["generator"] = {"reference": "pytorch/torchsnapshot @ /root/reference", "torch": torch.__version__}
    with open(os.path.join(OUT, f"{name}.json"), "w") as f:
        json.dump(dig, f, indent=1, sort_keys=False)
    print(f"{name}: {len(dig['manifest'])} entries, {len(dig['files'])} files")


for name, (build, knobs) in CASES.items():
    run(name, {"state": torchsnapshot.StateDict(**build("cpu"))}, knobs)

store = tempfile.NamedTemporaryFile(delete=False)
dist.init_process_group("gloo", init_method=f"file://{store.name}", rank=0, world_size=1)
for name, (_, knobs) in SHARDED_CASES.items():
    run(name, {"state": torchsnapshot.StateDict(**build_sharded(name, "cpu"))}, knobs)
dist.destroy_process_group()
