"""Config C1 of BASELINE.json: single-process Snapshot.take of a 1 GiB fp32 nn.Linear state_dict on the CPU
(no GPU involved).  Times this package's host path and, when /root/reference is present, the unmodified
reference, alternating A/B, warm, fresh directory per repetition."""
import json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torchsnapshot_b200 as B
ref = None
if os.path.isdir("/root/reference/torchsnapshot"):
    sys.path.insert(0, "/root/reference"); os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    import torchsnapshot as ref
torch.manual_seed(0)
model = torch.nn.Linear(16384, 16384)
nbytes = sum(p.numel() * 4 for p in model.parameters())
base = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="c1_", dir="/tmp")
res = {"ours_take": [], "ours_restore": [], "ref_take": [], "ref_restore": []}
for rep in range(6):
    for name, M in (("ours", B), ("ref", ref)):
        if M is None: continue
        d = os.path.join(base, f"{name}{rep}")
        t0 = time.perf_counter(); snap = M.Snapshot.take(d, {"model": model}); t1 = time.perf_counter()
        m2 = torch.nn.Linear(16384, 16384)
        t2 = time.perf_counter(); M.Snapshot(d).restore({"model": m2}); t3 = time.perf_counter()
        assert torch.equal(m2.weight, model.weight)
        if rep: res[f"{name}_take"].append(t1 - t0); res[f"{name}_restore"].append(t3 - t2)
        shutil.rmtree(d)
out = {k: {"median_ms": round(sorted(v)[len(v) // 2] * 1e3, 1), "GBps": round(nbytes / 1e9 / sorted(v)[len(v) // 2], 2)} for k, v in res.items() if v}
out["payload_bytes"] = nbytes; out["cpu_count"] = os.cpu_count(); out["dir"] = base
print(json.dumps(out))
