"""In-process sweep of the engine's host-side knobs on the C3 workload at N=1 (one box, one process): I/O worker
count and placement, ring geometry, O_DIRECT / fsync / arena-less / bounded-arena modes.  Each configuration gets
a fresh engine (N.reset_engines()), one warm-up take and `--reps` timed takes + one restore.  Prints one JSON line
per configuration; the caller keeps the log under profiles/."""
import argparse, json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import torchsnapshot_b200 as B
from torchsnapshot_b200 import _native as N, scheduler as S
from benchmarks import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--dir", default="/tmp")
ap.add_argument("--set", default="all")
args = ap.parse_args()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
f = tempfile.NamedTemporaryFile(delete=False); dist.init_process_group("gloo", init_method=f"file://{f.name}", rank=0, world_size=1)
local = W.build_llama_local(0, 1, dev)
payload = sum(t.numel() * t.element_size() for t, _, _ in local.values())
app = {"model": B.StateDict(**W.wrap_sharded(local, 0, dev))}
root = tempfile.mkdtemp(prefix="sweep_", dir=args.dir)
KEYS = ("TSNAP_B200_IO_THREADS", "TSNAP_B200_PINNED_SLOTS", "TSNAP_B200_PINNED_SLOT_BYTES", "TSNAP_B200_IO_PIN", "TSNAP_B200_ENGINE_FLAGS", "TSNAP_B200_HBM_STAGING_BYTES", "TSNAP_B200_NUMA", "TSNAP_B200_RING_NUMA")

def run(tag, env, do_async=False):
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    N.reset_engines()
    out = {"tag": tag, "env": env}
    try:
        B.Snapshot.take(os.path.join(root, "w"), app); shutil.rmtree(os.path.join(root, "w"))
        ts = []
        for r in range(args.reps):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            B.Snapshot.take(os.path.join(root, f"s{r}"), app)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
            st = (S.LAST_STATS.get("save") or [{}])[0]
            if r + 1 < args.reps: shutil.rmtree(os.path.join(root, f"s{r}"))
        out.update(take_ms=[round(x, 1) for x in ts], take_gbs=round(payload / 1e6 / (sum(ts) / len(ts)), 1), best_gbs=round(payload / 1e6 / min(ts), 1),
                   engine={k: round(st.get(k, 0), 1) for k in ("plan_ms", "kernel_ms", "copy_ms", "device_done_ms", "total_ms", "slot_wait_ms", "io_busy_ms", "io_queue_ms", "arena_bytes", "n_waves", "direct_bytes", "n_memcpy")},
                   phases={k: round(v, 1) for k, v in (S.LAST_STATS.get("take_phases_ms") or {}).items()})
        for t, _, _ in local.values(): t.zero_()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        B.Snapshot(os.path.join(root, f"s{args.reps - 1}")).restore(app)
        torch.cuda.synchronize(); rt = (time.perf_counter() - t0) * 1e3
        lt = (S.LAST_STATS.get("load") or [{}])[0]
        out.update(restore_ms=round(rt, 1), restore_gbs=round(payload / 1e6 / rt, 1), load_engine={k: round(lt.get(k, 0), 1) for k in ("plan_ms", "kernel_ms", "total_ms", "io_busy_ms", "arena_bytes", "n_waves", "direct_bytes")})
        shutil.rmtree(os.path.join(root, f"s{args.reps - 1}"))
        if do_async:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            p = B.Snapshot.async_take(os.path.join(root, "a"), app); t1 = time.perf_counter(); p.wait(); t2 = time.perf_counter()
            out.update(async_block_ms=round((t1 - t0) * 1e3, 1), async_total_ms=round((t2 - t0) * 1e3, 1))
            shutil.rmtree(os.path.join(root, "a"))
    except Exception as e:
        out["error"] = repr(e)[:300]
        shutil.rmtree(root, ignore_errors=True); os.makedirs(root, exist_ok=True)
    print(json.dumps(out), flush=True)

MiB = 1 << 20
base = {"TSNAP_B200_IO_THREADS": 16, "TSNAP_B200_PINNED_SLOTS": 64}
if args.set == "trace":
    # where do the occasional +50 ms takes come from?  12 traced takes, per-take landmarks of the engine timeline
    for k in KEYS: os.environ.pop(k, None)
    os.environ["TSNAP_B200_ENGINE_FLAGS"] = str(N.ENGINE_TRACE)
    N.reset_engines()
    B.Snapshot.take(os.path.join(root, "w"), app); shutil.rmtree(os.path.join(root, "w"))
    for r in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        B.Snapshot.take(os.path.join(root, f"t{r}"), app)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        st = (S.LAST_STATS.get("save") or [{}])[0]
        tr = (S.LAST_STATS.get("save_trace") or [[]])[0]
        def span(kind):
            rs = [x for x in tr if x["kind"] == kind]
            return [round(min(x["t0_ms"] for x in rs), 1), round(max(x["t1_ms"] for x in rs), 1)] if rs else None
        d2h = sorted((x for x in tr if x["kind"] == "d2h"), key=lambda x: x["t1_ms"])
        print(json.dumps({"take_ms": round(ms, 1), "phases": {k: round(v, 1) for k, v in (S.LAST_STATS.get("take_phases_ms") or {}).items() if v > 0.5},
                          "write_phases": {k: round(v, 1) for k, v in (S.LAST_STATS.get("write_phases_ms") or {}).items() if v > 0.5},
                          "engine": {k: round(st.get(k, 0), 1) for k in ("plan_ms", "kernel_ms", "device_done_ms", "copy_ms", "total_ms", "slot_wait_ms")},
                          "plan": span("plan"), "kernel": span("kernel"), "d2h": span("d2h"), "first_d2h_done": round(d2h[0]["t1_ms"], 1) if d2h else None,
                          "pwrite": span("pwrite"), "open": span("open"),
                          "slowest_d2h_gaps": sorted((round(b["t1_ms"] - a["t1_ms"], 1) for a, b in zip(d2h, d2h[1:])), reverse=True)[:3]}), flush=True)
        shutil.rmtree(os.path.join(root, f"t{r}"))
    shutil.rmtree(root, ignore_errors=True)
    dist.destroy_process_group()
    sys.exit(0)
run("base t16 s64x32", base, do_async=True)
if args.set in ("all", "threads"):
    for t in (8, 12, 20, 24, 32):
        run(f"t{t}", {**base, "TSNAP_B200_IO_THREADS": t})
    for t in (16, 24, 32):
        run(f"t{t} spread", {**base, "TSNAP_B200_IO_THREADS": t, "TSNAP_B200_IO_PIN": "spread"})
    run("t16 local", {**base, "TSNAP_B200_IO_PIN": "local"})
if args.set in ("all", "numa"):
    # placement of the pinned ring x placement of the workers; every combination twice (a fresh ring each time:
    # without explicit placement the ring lands wherever the allocating thread happens to run)
    for rep in range(2):
        for ring in ("none", "gpu", "interleave"):
            for pin in ("none", "node", "spread", "local"):
                if (ring, pin) in (("interleave", "local"),):
                    continue
                run(f"ring={ring} pin={pin} #{rep}", {**base, "TSNAP_B200_RING_NUMA": ring, "TSNAP_B200_IO_PIN": pin})
    for t in (12, 20, 24):
        run(f"ring=interleave pin=node t{t}", {**base, "TSNAP_B200_IO_THREADS": t, "TSNAP_B200_RING_NUMA": "interleave", "TSNAP_B200_IO_PIN": "node"})
    run("ring=interleave pin=node no_arena", {**base, "TSNAP_B200_RING_NUMA": "interleave", "TSNAP_B200_IO_PIN": "node", "TSNAP_B200_ENGINE_FLAGS": N.ENGINE_NO_ARENA})
if args.set in ("all", "ring"):
    for sb, n in ((8, 256), (16, 128), (64, 32), (32, 32), (32, 128)):
        run(f"slots {n}x{sb}MiB", {**base, "TSNAP_B200_PINNED_SLOTS": n, "TSNAP_B200_PINNED_SLOT_BYTES": sb * MiB})
    run("slots 256x8MiB t24 spread", {"TSNAP_B200_IO_THREADS": 24, "TSNAP_B200_IO_PIN": "spread", "TSNAP_B200_PINNED_SLOTS": 256, "TSNAP_B200_PINNED_SLOT_BYTES": 8 * MiB})
if args.set in ("all", "modes"):
    run("no_arena (direct D2H from live tensors)", {**base, "TSNAP_B200_ENGINE_FLAGS": N.ENGINE_NO_ARENA}, do_async=True)
    run("arena cap 2 GiB", {**base, "TSNAP_B200_HBM_STAGING_BYTES": 2 << 30}, do_async=True)
    run("arena cap 1 GiB", {**base, "TSNAP_B200_HBM_STAGING_BYTES": 1 << 30}, do_async=True)
    run("arena cap 4 GiB", {**base, "TSNAP_B200_HBM_STAGING_BYTES": 4 << 30}, do_async=True)
    run("O_DIRECT", {**base, "TSNAP_B200_ENGINE_FLAGS": N.ENGINE_ODIRECT})
    run("fsync (durable)", {**base, "TSNAP_B200_ENGINE_FLAGS": N.ENGINE_FSYNC})
    run("O_DIRECT + fsync (durable)", {**base, "TSNAP_B200_ENGINE_FLAGS": N.ENGINE_ODIRECT | N.ENGINE_FSYNC})
shutil.rmtree(root, ignore_errors=True)
dist.destroy_process_group()
