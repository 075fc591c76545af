"""How close is oracle/ref_port.RefPipeline (the timed "port" of the reference's execution pipeline) to the real
reference?  Same CPU tensors, same directory type, alternating A/B, warm.  Build container only (needs /root/reference)."""
import json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference"); os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
import torch, torchsnapshot as ref
from oracle.ref_port import RefPipeline
torch.manual_seed(0)
# Llama-like mix scaled down: 2 x 128 MiB + 24 x 24 MiB + 48 x 4 MiB + 48 x 8 KiB  ~= 1 GiB, bf16
shapes = [(16384, 4096)] * 2 + [(3072, 4096)] * 24 + [(512, 4096)] * 48 + [(4096,)] * 48
tensors = {f"t{i}": torch.randn(s).to(torch.bfloat16) for i, s in enumerate(shapes)}
nbytes = sum(t.numel() * 2 for t in tensors.values())
base = tempfile.mkdtemp(prefix="fidelity_", dir="/tmp")
res = {"reference_take": [], "port_save": [], "reference_restore": [], "port_load": []}
for rep in range(6):
    d = os.path.join(base, f"r{rep}")
    t0 = time.perf_counter(); snap = ref.Snapshot.take(d, {"m": ref.StateDict(**tensors)}); t1 = time.perf_counter()
    out = ref.StateDict(**{k: torch.zeros_like(v) for k, v in tensors.items()})
    t2 = time.perf_counter(); ref.Snapshot(d).restore({"m": out}); t3 = time.perf_counter()
    shutil.rmtree(d)
    d = os.path.join(base, f"p{rep}")
    t4 = time.perf_counter(); pipe = RefPipeline(d); idx = pipe.save(tensors); t5 = time.perf_counter()
    out2 = {k: torch.zeros_like(v) for k, v in tensors.items()}
    t6 = time.perf_counter(); pipe.load(idx, out2); t7 = time.perf_counter()
    assert all(torch.equal(out2[k], tensors[k]) and torch.equal(out[k], tensors[k]) for k in tensors)
    shutil.rmtree(d)
    if rep:
        res["reference_take"].append(t1 - t0); res["reference_restore"].append(t3 - t2); res["port_save"].append(t5 - t4); res["port_load"].append(t7 - t6)
shutil.rmtree(base, ignore_errors=True)
med = lambda v: sorted(v)[len(v) // 2]
print(json.dumps({k: {"median_ms": round(med(v) * 1e3, 1), "GBps": round(nbytes / 1e9 / med(v), 2)} for k, v in res.items()} | {"payload_bytes": nbytes, "cpu_count": os.cpu_count()}))
