"""Config C4 of BASELINE.json: async_take during a running DDP training loop (GPT-2-medium-sized transformer,
fp32 params + AdamW).  Reports step time without snapshots, the blocking window of async_take, the step-time
inflation while a snapshot drains in the background, and overlap % = 1 - (extra step time) / (drain time)."""
import json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
import torchsnapshot_b200 as B

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if world > 1: dist.init_process_group("nccl", device_id=dev)
else:
    f = tempfile.NamedTemporaryFile(delete=False); dist.init_process_group("gloo", init_method=f"file://{f.name}", rank=0, world_size=1)
torch.manual_seed(0)
d, layers, vocab, ctx = 1024, 24, 50257, 512
class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.emb = torch.nn.Embedding(vocab, d); self.pos = torch.nn.Embedding(1024, d)
        layer = torch.nn.TransformerEncoderLayer(d, 16, 4 * d, dropout=0.0, batch_first=True, norm_first=True)
        self.blocks = torch.nn.TransformerEncoder(layer, layers)
        self.ln = torch.nn.LayerNorm(d)
    def forward(self, x):
        h = self.emb(x) + self.pos(torch.arange(x.shape[1], device=x.device))
        return self.ln(self.blocks(h)) @ self.emb.weight.t()
model = Net().to(dev)
ddp = DDP(model, device_ids=[local]) if world > 1 else model
opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4)
nparams = sum(p.numel() for p in model.parameters())
def step():
    x = torch.randint(0, vocab, (4, ctx), device=dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(ddp(x).float().view(-1, vocab), x.view(-1))
    loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
def timed_steps(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): step()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for _ in range(3): step()
app = {"model": ddp, "optim": opt}
kw = {"replicated": ["optim/**"]} if world > 1 else {}
box = [tempfile.mkdtemp(prefix="c4_") if rank == 0 else None]
if world > 1: dist.broadcast_object_list(box, src=0)
root = box[0]
B.Snapshot.take(os.path.join(root, "warm"), app, **kw)  # pins the ring, grows the arena
base = timed_steps(10)
res = []
STEPS_DURING = 6  # fixed count on every rank: the loop must not depend on rank-local completion (DDP collectives)
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pending = B.Snapshot.async_take(os.path.join(root, f"s{k}"), app, **kw)
    t1 = time.perf_counter()
    for _ in range(STEPS_DURING): step()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    still_draining = not pending.done()
    pending.wait()
    t3 = time.perf_counter()
    res.append({"blocked_ms": (t1 - t0) * 1e3, "steps_ms_total": (t2 - t1) * 1e3, "step_ms_during_drain": (t2 - t1) * 1e3 / STEPS_DURING,
                "wait_after_steps_ms": (t3 - t2) * 1e3, "still_draining_after_steps": float(still_draining)})
payload = sum(p.numel() * 4 for p in model.parameters()) * 3
if rank == 0:
    r = res[-1]
    extra = max(0.0, r["step_ms_during_drain"] - base * 1e3) * STEPS_DURING
    drain_ms = r["steps_ms_total"] + r["wait_after_steps_ms"]
    print(json.dumps({"config": "C4", "world": world, "params": nparams, "state_bytes_model+adam": payload, "step_ms": round(base * 1e3, 2),
                      "async_take": [{k: round(v, 2) for k, v in x.items()} for x in res],
                      "overlap_pct": round(100 * (1 - extra / drain_ms), 1)}), flush=True)
if world > 1: dist.barrier(device_ids=[local])
if rank == 0: shutil.rmtree(root, ignore_errors=True)
dist.destroy_process_group()
