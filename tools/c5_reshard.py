"""Config C5 of BASELINE.json at full size: a 16 GB row-wise ShardedTensor (31_250_000 x 128 fp32) saved at
world_size=8 and restored at world_size=4 (reshard-on-load).  Phase is chosen by --phase save|restore; both run under
torchrun on the same box against the same directory.  Correctness is checked with a closed-form content function."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata
import torchsnapshot_b200 as B

ROWS, COLS = 31_250_000, 128

def content(lo, n, dev):
    r = torch.arange(lo, lo + n, device=dev, dtype=torch.int64).unsqueeze(1)
    c = torch.arange(COLS, device=dev, dtype=torch.int64).unsqueeze(0)
    return ((r * 131 + c * 7) % 65521).to(torch.float32)

ap = argparse.ArgumentParser(); ap.add_argument("--phase", required=True); ap.add_argument("--dir", required=True); args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
split = -(-ROWS // world); lo = rank * split; n = min(split, ROWS - lo)
def sharded(t):
    return ShardedTensor._init_from_local_shards([Shard(tensor=t, metadata=ShardMetadata(shard_offsets=[lo, 0], shard_sizes=[n, COLS], placement=f"rank:{rank}/cuda:{local}"))], (ROWS, COLS))
def tmax(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t)
if args.phase == "save":
    local_t = torch.empty(n, COLS, device=dev)
    for a in range(0, n, 1 << 20):
        local_t[a:a + (1 << 20)] = content(lo + a, min(1 << 20, n - a), dev)
    app = {"emb": B.StateDict(table=sharded(local_t))}
    B.Snapshot.take(os.path.join(args.dir, "warm"), app)  # warm-up: pins the ring, grows the arena
    dist.barrier(device_ids=[local]); torch.cuda.synchronize(); t0 = time.perf_counter()
    snap = B.Snapshot.take(os.path.join(args.dir, "snap"), app)
    torch.cuda.synchronize(); dt = tmax(time.perf_counter() - t0)
    if rank == 0:
        man = snap.get_manifest()
        pieces = [len(man[f"{r}/emb/table"].shards) for r in range(world)]
        print(json.dumps({"phase": "save", "world": world, "take_s": round(dt, 3), "GBps": round(ROWS * COLS * 4 / 1e9 / dt, 1), "pieces_per_rank": pieces}), flush=True)
else:
    local_t = torch.zeros(n, COLS, device=dev)
    app = {"emb": B.StateDict(table=sharded(local_t))}
    snap = B.Snapshot(os.path.join(args.dir, "snap"))
    snap.restore(app)  # warm-up
    local_t.zero_()
    dist.barrier(device_ids=[local]); torch.cuda.synchronize(); t0 = time.perf_counter()
    snap.restore(app)
    torch.cuda.synchronize(); dt = tmax(time.perf_counter() - t0)
    ok = True
    for a in range(0, n, 1 << 20):
        m = min(1 << 20, n - a)
        ok = ok and bool(torch.equal(local_t[a:a + m], content(lo + a, m, dev)))
    flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"phase": "restore", "world": world, "restore_s": round(dt, 3), "GBps": round(ROWS * COLS * 4 / 1e9 / dt, 1), "verified_all_ranks": bool(flag.item())}), flush=True)
dist.barrier(device_ids=[local]); dist.destroy_process_group()
