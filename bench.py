#!/usr/bin/env python3
"""Headline benchmark: checkpoint GB/s and Snapshot.take() blocking ms for an FSDP-layout Llama-3-8B bf16
sharded state dict (BASELINE.json config C3), one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's engine
    python bench.py --impl reference --gpus 1 --steps K --warmup W   # the reference's CPU/asyncio path (oracle port)

One "step" = one ``Snapshot.take`` of the whole (per-rank sharded) state dict into a fresh directory on the
local filesystem.  The JSON line reports
  e2e.value  : GB/s of Snapshot.take through the public API (wall clock; D2H copies + file writes inside)
  value      : GB/s of the device-side drain alone (pack kernels + D2H into pinned host memory through the
               C-ABI stager seam), timed with CUDA events on the engine's streams
  roofline   : the pack kernel against the measured HBM copy peak (algorithmic traffic 2 x payload)
  cpu_baseline: the reference's pipeline (oracle/ref_port.RefPipeline) timed on this box's host cores
plus blocking ms of async_take, restore GB/s and the D2H link fraction.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

# ---- workload: Llama-3-8B parameter shapes (vocab 128256, dim 4096, 32 layers, ffn 14336, 8 KV heads) ----
def llama3_8b_shapes():
    dim, ffn, vocab, layers, kv = 4096, 14336, 128256, 32, 1024
    shapes = [("tok_embeddings.weight", (vocab, dim))]
    for i in range(layers):
        p = f"layers.{i}."
        shapes += [
            (p + "attention.wq.weight", (dim, dim)),
            (p + "attention.wk.weight", (kv, dim)),
            (p + "attention.wv.weight", (kv, dim)),
            (p + "attention.wo.weight", (dim, dim)),
            (p + "feed_forward.w1.weight", (ffn, dim)),
            (p + "feed_forward.w2.weight", (dim, ffn)),
            (p + "feed_forward.w3.weight", (ffn, dim)),
            (p + "attention_norm.weight", (dim,)),
            (p + "ffn_norm.weight", (dim,)),
        ]
    shapes += [("norm.weight", (dim,)), ("output.weight", (vocab, dim))]
    return shapes


def local_rows(rows: int, rank: int, world: int):
    """dim-0 chunk of ChunkShardingSpec / FSDP sharded state dicts: ceil split, ragged tail."""
    split = -(-rows // world)
    lo = min(rank * split, rows)
    return lo, max(0, min(split, rows - lo))


def build_local_tensors(rank: int, world: int, device: torch.device, seed: int = 42):
    """{name: (local bf16 tensor, global shape, row offset)} — synthetic weights, random init."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + rank)
    out = {}
    for name, shape in llama3_8b_shapes():
        lo, n = local_rows(shape[0], rank, world)
        local = torch.empty((n,) + tuple(shape[1:]), dtype=torch.bfloat16, device=device)
        if local.numel():
            local.normal_(generator=gen)
        out[name] = (local, shape, lo)
    return out


def wrap_sharded(local, rank: int, device: torch.device):
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata

    state = {}
    for name, (t, shape, lo) in local.items():
        off = [lo] + [0] * (len(shape) - 1)
        md = ShardMetadata(shard_offsets=off, shard_sizes=list(t.shape), placement=f"rank:{rank}/{device}")
        state[name] = ShardedTensor._init_from_local_shards([Shard(tensor=t, metadata=md)], tuple(shape))
    return state


def wrap_dtensor(local, world: int, device: torch.device):
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor, Shard as ShardPlacement

    mesh = init_device_mesh(device.type, (world,))
    return {name: DTensor.from_local(t, mesh, [ShardPlacement(0)], run_check=False, shape=torch.Size(shape), stride=torch.empty(shape, device="meta").stride())
            for name, (t, shape, lo) in local.items()}


# ---- clocks sampling (B200_PROFILING.md recipe) ---------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int) -> None:
        self.index = index
        self.proc = None
        self.lines = []

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self) -> None:
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---- helpers -------------------------------------------------------------------------------------------
def dist_max(x: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dist_sum(x: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(device) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            dist.barrier(device_ids=[device.index])
        else:
            dist.barrier()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def rm_tree(path: str) -> None:
    shutil.rmtree(path, ignore_errors=True)


def sample_for_cpu(local, limit_bytes: int):
    """Bounded sample of the workload for the reference arm: leading tensors up to `limit_bytes`."""
    out, total = {}, 0
    for name, (t, _, _) in local.items():
        nb = t.numel() * t.element_size()
        if total + nb > limit_bytes and out:
            continue
        out[name] = t
        total += nb
    return out, total


# ---- reference arm ---------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int, device: torch.device, base_dir: str) -> None:
    if rank != 0:
        return
    from oracle.ref_port import RefPipeline

    local = build_local_tensors(0, world, device)
    sample, nbytes = sample_for_cpu(local, args.ref_sample_gib << 30)
    sample_desc = f"{len(sample)} of {len(local)} tensors of rank 0's shard ({nbytes / 1e9:.2f} GB of {sum(t.numel() * 2 for t, _, _ in local.values()) / 1e9:.2f} GB)"
    times = []
    restore_times = []
    for step in range(args.warmup + args.steps):
        d = os.path.join(base_dir, f"ref{step}")
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        pipe = RefPipeline(d)
        index = pipe.save(sample)
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
            if len(restore_times) < 2:
                out = {k: torch.zeros_like(v) for k, v in sample.items()}
                torch.cuda.synchronize(device)
                t0 = time.perf_counter()
                pipe.load(index, out)
                torch.cuda.synchronize(device)
                restore_times.append(time.perf_counter() - t0)
        rm_tree(d)
    ms = 1e3 * sum(times) / len(times)
    gbs = nbytes / 1e9 / (ms / 1e3)
    cores = RefPipeline.CPU_THREADS + RefPipeline.IO_CONCURRENCY + 1
    line = {
        "impl": "reference",
        "metric": "checkpoint_save_GBps",
        "value": gbs,
        "unit": "GB/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "bf16 (byte copy)",
        "data": "synthetic",
        "config": {"workload": "FSDP-layout Llama-3-8B bf16 sharded state_dict, Snapshot.take to local fs", "world_size": world,
                   "note": "reference's CPU/asyncio pipeline restated (oracle/ref_port.RefPipeline): pageable tensor.to('cpu') in 4 threads, per-member D2D + blocking .cpu() per GPU slab, <=16 file writes"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "host_cores_available": os.cpu_count(), "kind": "port", "sample": sample_desc,
                         "restore_gbs": nbytes / 1e9 / (sum(restore_times) / len(restore_times)) if restore_times else None},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---- our arm -------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, device: torch.device, base_dir: str) -> None:
    import torchsnapshot_b200 as B
    from torchsnapshot_b200 import _native as N
    from torchsnapshot_b200 import scheduler as S

    local = build_local_tensors(rank, world, device)
    payload_local = sum(t.numel() * t.element_size() for t, _, _ in local.values())
    payload_total = dist_sum(float(payload_local), device)
    if args.layout == "dtensor":
        state = wrap_dtensor(local, world, device)
    elif args.layout == "plain":
        state = {k: t for k, (t, _, _) in local.items()}
    else:
        state = wrap_sharded(local, rank, device)
    app_state = {"model": B.StateDict(**state)}
    eng = B.get_engine(device.index)

    def snap_dir(tag):
        return os.path.join(base_dir, tag)

    def cleanup(tag):
        barrier(device)
        if rank == 0:
            rm_tree(snap_dir(tag))
        barrier(device)

    # -- warm-up (pins the ring, grows the HBM arena, warms the page cache paths) --
    for w in range(args.warmup):
        B.Snapshot.take(snap_dir(f"warm{w}"), app_state)
        cleanup(f"warm{w}")

    # -- timed e2e steps: Snapshot.take through the public API --
    launches0 = eng.stats()["kernels_launched"]
    sampler = ClockSampler(device.index)
    if rank == 0:
        sampler.start()
    step_ms, kernel_ms, kernel_bulk_ms, copy_ms, plan_ms, table_bytes = [], [], [], [], [], []
    for k in range(args.steps):
        barrier(device)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        B.Snapshot.take(snap_dir(f"step{k}"), app_state)
        torch.cuda.synchronize(device)
        dt = (time.perf_counter() - t0) * 1e3
        step_ms.append(dist_max(dt, device))
        st = (S.LAST_STATS.get("save") or [{}])[0]
        kernel_ms.append(st.get("kernel_ms", 0.0))
        kernel_bulk_ms.append(st.get("kernel_bulk_ms", 0.0))
        copy_ms.append(st.get("copy_ms", 0.0))
        plan_ms.append(st.get("plan_ms", 0.0))
        table_bytes.append(st.get("table_h2d_bytes", 0))
        last_stats = st
        if k + 1 < args.steps:
            cleanup(f"step{k}")
    clocks = sampler.stop() if rank == 0 else {}
    e2e_ms = sum(step_ms) / len(step_ms)
    e2e_gbs = payload_total / 1e9 / (e2e_ms / 1e3)
    keep_tag = f"step{args.steps - 1}"  # kept for the restore measurement
    if args.only_e2e:
        cleanup(keep_tag)
        if rank == 0:
            emit({"only_e2e": True, "take_phases_ms": S.LAST_STATS.get("take_phases_ms"), "write_phases_ms": S.LAST_STATS.get("write_phases_ms"), "e2e_gbs": e2e_gbs, "e2e_ms": e2e_ms, "steps_ms": step_ms, "engine_step": last_stats,
                              "io_threads": os.environ.get("TSNAP_B200_IO_THREADS"), "slots": os.environ.get("TSNAP_B200_PINNED_SLOTS"),
                              "slot_bytes": os.environ.get("TSNAP_B200_PINNED_SLOT_BYTES")})
        return

    # -- restore (same snapshot) --
    restore_ms = []
    for _ in range(min(args.steps, 3)):
        for t, _, _ in local.values():
            t.zero_()
        barrier(device)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        B.Snapshot(snap_dir(keep_tag)).restore(app_state)
        torch.cuda.synchronize(device)
        restore_ms.append(dist_max((time.perf_counter() - t0) * 1e3, device))
    load_stats = (S.LAST_STATS.get("load") or [{}])[0]
    # the restored state must be the saved one: per-rank checksum of the first tensors against a regeneration
    regen = build_local_tensors(rank, world, device)
    ok = all(torch.equal(local[n][0], regen[n][0]) for n in list(local)[:12])
    del regen
    cleanup(keep_tag)

    # -- async_take blocking window --
    block_ms, async_total_ms = [], []
    for k in range(min(args.steps, 3)):
        barrier(device)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        pending = B.Snapshot.async_take(snap_dir(f"async{k}"), app_state)
        t1 = time.perf_counter()
        pending.wait()
        t2 = time.perf_counter()
        block_ms.append(dist_max((t1 - t0) * 1e3, device))
        async_total_ms.append(dist_max((t2 - t0) * 1e3, device))
        blocking_stats = (S.LAST_STATS.get("save") or [{}])[0]
        cleanup(f"async{k}")

    # -- device-side drain alone (value): pack kernels + D2H into pinned memory via the stager seam --
    descs, keep = [], []
    off = 0
    for t, _, _ in local.values():
        if t.numel():
            descs.append(N.save_desc(t, off))
            keep.append(t)
        off += t.numel() * t.element_size()
    dev_ms = []
    stream = torch.cuda.current_stream(device).cuda_stream
    for k in range(args.warmup + args.steps):
        barrier(device)
        torch.cuda.synchronize(device)
        sb = eng.stage(descs, off, stream=stream, keepalive=keep)
        sb.wait()
        h = sb.stats()
        sb.release()
        if k >= args.warmup:
            dev_ms.append((dist_max(h["kernel_ms"] + h["copy_ms"], device), dist_max(h["kernel_ms"], device), dist_max(h["copy_ms"], device), h["kernel_bulk_ms"], h["n_tiles_bulk"], h["n_tiles_lsu"]))
    eng.trim()
    value_ms = sum(x[0] for x in dev_ms) / len(dev_ms)
    value_gbs = payload_total / 1e9 / (value_ms / 1e3)
    pack_ms = sum(x[1] for x in dev_ms) / len(dev_ms)
    d2h_ms = sum(x[2] for x in dev_ms) / len(dev_ms)
    bulk_ms = sum(x[3] for x in dev_ms) / len(dev_ms)
    launches = eng.stats()["kernels_launched"] - launches0

    # -- cpu baseline: the reference's pipeline on a bounded sample (rank 0, N=1) --
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        from oracle.ref_port import RefPipeline

        sample, nbytes = sample_for_cpu(local, args.ref_sample_gib << 30)
        times = []
        for rep in range(3):
            d = snap_dir(f"cpu{rep}")
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            RefPipeline(d).save(sample)
            torch.cuda.synchronize(device)
            if rep:
                times.append(time.perf_counter() - t0)
            rm_tree(d)
        cpu_baseline = {
            "value": nbytes / 1e9 / (sum(times) / len(times)),
            "unit": "GB/s",
            "cores": RefPipeline.CPU_THREADS + RefPipeline.IO_CONCURRENCY + 1,
            "host_cores_available": os.cpu_count(),
            "kind": "port",
            "sample": f"{len(sample)} of {len(local)} tensors ({nbytes / 1e9:.2f} GB of {payload_local / 1e9:.2f} GB), 2 timed passes after 1 warm-up",
        }

    if rank != 0:
        return
    traffic = args.ncu_traffic_bytes
    if traffic is None:
        # per-launch DRAM traffic of the same kernel on the same workload from the committed ncu capture
        try:
            with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
                cap = json.load(f)
            if cap.get("payload_bytes_per_rank") == int(payload_local):
                traffic = cap["traffic_bytes"]
        except Exception:
            traffic = None
    peaks, peaks_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # dominant kernel = the bulk (TMA) pack kernel; per launch it moves this rank's payload twice (read + write)
    bulk_bytes = 2.0 * payload_local
    achieved = bulk_bytes / 1e9 / (bulk_ms / 1e3) if bulk_ms > 0 else 0.0
    link_peak = args.link_peak_gbs
    line = {
        "metric": "checkpoint_save_GBps",
        "value": value_gbs,
        "unit": "GB/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": value_ms,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "bf16 (byte copy)",
        "data": "synthetic",
        "config": {
            "workload": "FSDP-layout Llama-3-8B bf16 sharded state_dict (291 tensors, 16.06 GB total, dim-0 sharded over n_gpus), Snapshot.take to local fs",
            "layout": args.layout,
            "payload_bytes_total": int(payload_total),
            "payload_bytes_per_rank": int(payload_local),
            "target_dir": base_dir,
            "l2": "inputs (>=2 GB per rank) exceed the 126 MB L2; every step writes a fresh directory",
            "value_definition": "device-side drain: pack kernels + D2H into pinned host memory (C-ABI stager seam), CUDA-event timed, max over ranks",
        },
        "e2e": {
            "value": e2e_gbs,
            "unit": "GB/s",
            "ms_per_step": e2e_ms,
            "h2d_bytes_per_step": int(sum(table_bytes) / max(1, len(table_bytes))),
            "d2h_bytes_per_step": int(payload_local),
            "definition": "Snapshot.take(path, app_state) wall clock incl. planning collectives, D2H, file writes, metadata commit; max over ranks",
        },
        "take_blocking_ms": {"async_take_returns_ms": sum(block_ms) / len(block_ms), "async_total_ms": sum(async_total_ms) / len(async_total_ms),
                             "engine_device_done_ms": blocking_stats.get("device_done_ms"), "sync_take_ms": e2e_ms},
        "restore": {"value": payload_total / 1e9 / (sum(restore_ms) / len(restore_ms) / 1e3), "unit": "GB/s", "ms": sum(restore_ms) / len(restore_ms),
                    "verified": bool(ok), "scatter_kernel_ms": load_stats.get("kernel_ms")},
        "roofline": {
            "bound": "hbm",
            "kernel": "tsnap_bulk_copy_kernel (pack)",
            "achieved": achieved,
            "peak": hbm_peak,
            "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks_src})",
            "unit": "GB/s",
            "frac": achieved / hbm_peak if hbm_peak else None,
            "algorithmic_bytes_per_launch": bulk_bytes,
            "launch_ms": bulk_ms,
            "traffic": traffic,
        },
        "link": {"achieved": payload_local / 1e9 / (d2h_ms / 1e3) if d2h_ms else None, "peak": link_peak, "unit": "GB/s",
                 "frac": (payload_local / 1e9 / (d2h_ms / 1e3)) / link_peak if d2h_ms else None,
                 "peak_source": "pinned cudaMemcpyAsync D2H measured on this pool (profiles/r01_box_probe.json)", "d2h_ms": d2h_ms, "pack_ms": pack_ms},
        "take_phases_ms": {k: round(v, 2) for k, v in (S.LAST_STATS.get("take_phases_ms") or {}).items()},
        "write_phases_ms": {k: round(v, 2) for k, v in (S.LAST_STATS.get("write_phases_ms") or {}).items()},
        "engine_step": {k: last_stats.get(k) for k in ("plan_ms", "kernel_ms", "copy_ms", "device_done_ms", "total_ms", "n_files", "n_members", "n_tiles_bulk", "n_tiles_lsu", "n_kernel_launches")},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "host": {"cpu_count": os.cpu_count(), "ranks_on_host": world, "cores_per_rank": (os.cpu_count() or 0) // max(1, world),
                 "engine_io_threads_per_rank": int(os.environ.get("TSNAP_B200_IO_THREADS", max(2, 16 // max(1, world))))},
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The single JSON line goes to the process's original stdout; everything else (NCCL banners, warnings
    printed by libraries) was re-routed to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main() -> None:
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # NCCL prints its version banner on fd 1
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--layout", choices=["sharded", "dtensor", "plain"], default="sharded")
    ap.add_argument("--dir", default=None, help="target directory (default: $TSNAP_BENCH_DIR or a temp dir under /tmp)")
    ap.add_argument("--ref-sample-gib", type=int, default=4, help="bounded sample of the workload for the reference arm")
    ap.add_argument("--link-peak-gbs", type=float, default=57.0)
    ap.add_argument("--ncu-traffic-bytes", type=float, default=None)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--only-e2e", action="store_true", help="tuning aid: run only the timed Snapshot.take steps and print a short line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback for device tensors")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if args.impl == "reference" and rank != 0:
        return
    store_file = None
    if world > 1 and args.impl == "ours":
        dist.init_process_group("nccl", device_id=device)
    else:
        store_file = tempfile.NamedTemporaryFile(delete=False)
        dist.init_process_group("gloo", init_method=f"file://{store_file.name}", rank=0, world_size=1)
    base = args.dir or os.environ.get("TSNAP_BENCH_DIR")
    made = False
    if base is None:
        box = [tempfile.mkdtemp(prefix="tsnap_bench_", dir="/tmp") if rank == 0 else None]
        if world > 1 and args.impl == "ours":
            dist.broadcast_object_list(box, src=0)
        base = box[0]
        made = True
    os.makedirs(base, exist_ok=True)
    try:
        if args.impl == "reference":
            run_reference(args, rank, world, device, base)
        else:
            run_ours(args, rank, world, device, base)
    finally:
        if dist.is_initialized():
            try:
                barrier(device)
            except Exception:
                pass
        if rank == 0 and made:
            rm_tree(base)
        if dist.is_initialized():
            dist.destroy_process_group()
        if store_file is not None:
            try:
                os.unlink(store_file.name)
            except OSError:
                pass


if __name__ == "__main__":
    main()
