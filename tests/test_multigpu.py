"""Multi-GPU functional parity (NCCL): launches tools/mgpu_check.py under torchrun on every GPU of the box (2..8) —
DDP auto-replication + partitioned writes, Adam state under a replicated glob, per-rank state, ShardedTensor, sync and
async commit (TCPStore two-phase), restore into perturbed copies, read_object resharding — and keeps the log.
Mirrors T:tests/test_ddp.py:49-138 and T:tests/gpu_tests/ (which need >= 2 GPUs as well)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_multi_gpu_functional_check_under_torchrun():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29571",
           os.path.join(ROOT, "tools", "mgpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"mgpu_check_n{n}.log"), "w") as f:
        f.write(res.stdout + "\n--- stderr ---\n" + res.stderr[-20000:])
    assert res.returncode == 0, res.stderr[-3000:]
    assert res.stdout.count("multi-GPU check OK") == n, res.stdout[-2000:]  # (ranks' lines may interleave)
