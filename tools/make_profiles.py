"""Regenerates the markdown summaries under profiles/ from the raw artefacts in profiles/r02_raw/ (bench JSON lines,
kernel-case lines, timeline summaries).  Run after copying new raw files in; nothing here measures anything."""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = os.path.join(ROOT, "profiles", "r02_raw")

def load(name):
    try:
        return json.load(open(os.path.join(RAW, name)))
    except Exception:
        return None

def f(x, nd=1):
    return "—" if x is None else (f"{x:.{nd}f}" if isinstance(x, (int, float)) else str(x))

# ---- bench summary ---------------------------------------------------------------------------------------------------
out = ["# r02 — bench.py results on the B200 pool (both arms run the same code; reference = unmodified torchsnapshot from oracle/_ref)", "",
       "Raw JSON lines: `profiles/r02_raw/<config>_n<N>_{ours,reference}.json`.  Hosts: 2 x Xeon 8562Y+ (128 hw threads, 2 NUMA nodes), 2 TB RAM, PCIe Gen5 x16 per GPU;",
       "target = `/tmp` (overlayfs on one NVMe), page-cache (\"returned\") semantics on both arms.  `value` = Σ payload / max-over-ranks wall of `Snapshot.take`.", ""]
out += ["## C3 — FSDP-layout Llama-3-8B bf16, 16.06 GB total (strong scaling: per-rank payload = 16.06/N GB)", "",
        "| N | arm | take GB/s | take ms | restore GB/s | async_take blocks (ms) | link Σ D2H (GB/s) | host write peak (GB/s) | take / min(link, sink) | restore / min(H2D, source) | pack kernel frac of HBM peak |", "|---|---|---|---|---|---|---|---|---|---|---|"]
for n in (1, 2, 4, 8):
    for arm in ("ours", "reference"):
        d = load(f"c3_n{n}_{arm}.json")
        if not d:
            continue
        r = d.get("e2e_roofline") or {}
        out.append(f"| {n} | {arm} | **{f(d['value'])}** | {f(d['ms_per_step'], 0)} | {f(d['restore']['value'])} | {f(d['take_blocking_ms']['async_take_returns_ms'], 0)} | {f(r.get('link_d2h_gbs_all_ranks'))} | {f(r.get('host_write_peak_gbs'))} | {f(r.get('frac'), 3)} | {f(r.get('restore_frac'), 3)} | {f((d.get('roofline') or {}).get('frac'), 3)} |")
out += ["", "Speed-up of `Snapshot.take` over the unmodified reference on the same box, same ranks, same app_state, full payload:", ""]
for n in (1, 2, 4, 8):
    a, b = load(f"c3_n{n}_ours.json"), load(f"c3_n{n}_reference.json")
    if a and b:
        out.append(f"* N={n}: take {a['value'] / b['value']:.1f}x, restore {a['restore']['value'] / b['restore']['value']:.1f}x, async_take blocking {b['take_blocking_ms']['async_take_returns_ms'] / max(a['take_blocking_ms']['async_take_returns_ms'], 1e-9):.0f}x shorter")
d = load("c3_n1_ours.json")
if d:
    out += ["", f"N=1 detail: steps (ms) {d.get('steps_ms')}; engine of the last step {json.dumps(d.get('engine_step'))};",
            f"cpu_baseline {json.dumps(d.get('cpu_baseline'))}; timeline {json.dumps(d.get('timeline'))}."]
c1 = [(tag, load(f"c1_{tag}_ours.json"), load(f"c1_{tag}_reference.json")) for tag in ("buildbox", "gpubox")]
if any(a and b for _, a, b in c1):
    out += ["", "## C1 — single-process 1 GiB fp32 nn.Linear on the CPU (no GPU on the path; the host engine: planner + native pwrite/pread workers)", "",
            "| host | arm | take GB/s | take ms (each) | restore GB/s | cores busy |", "|---|---|---|---|---|---|"]
    for tag, a, b in c1:
        for arm, d in (("ours", a), ("reference", b)):
            if d:
                out.append(f"| {tag} ({d['host']['cpu_count']} hw threads) | {arm} | **{f(d['value'], 2)}** | {d['steps_ms']} | {f(d['restore']['value'], 2)} | {d['host_cpu_during_take']['cores_busy_avg']} |")
out += ["", "## C2 — DDP ResNet-50 + Adam, replicated=['**'] (0.31 GB, ≈800 tensors, 0-d scalars, CPU `step` slab chain)", "",
        "| N | arm | take ms | take GB/s | restore ms | control plane / engine job (ms) | collectives per take | LSU + bulk kernel ms |", "|---|---|---|---|---|---|---|---|"]
for n in (2, 8):
    for arm in ("ours", "reference"):
        d = load(f"c2_n{n}_{arm}.json")
        if not d:
            continue
        c = d.get("control_vs_data_ms") or {}
        e = d.get("engine_step") or {}
        out.append(f"| {n} | {arm} | **{f(d['ms_per_step'])}** | {f(d['value'], 2)} | {f(d['restore']['ms'], 0)} | {f(c.get('control_plane_ms'))} / {f(c.get('engine_job_total_ms'))} | {f(c.get('collectives_per_take'), 0)} | {f(e.get('kernel_lsu_ms'), 4)} + {f(e.get('kernel_bulk_ms'), 4)} |")
out += ["", "The LSU kernel moves the unaligned slab members here (a few KB per rank after the 8-way partition) — launch-latency, not bandwidth, territory; its roofline numbers come from `r02_kernel_cases.md`.",
        "Restore uses the read-once path: each replicated range is read from storage by one rank and broadcast over NVLink (`mgpu_check_n8.log`: 19.12 MB read over 8 ranks for 19.13 MB on disk).", ""]
out += ["## C4 — GPT-2-medium DDP training loop + async_take (4.26 GB of fp32 state, replicated)", "",
        "| N | arm | async_take blocks (ms) | train step (ms) | step during drain (ms) | overlap % | training time lost per snapshot (ms) | sync take (ms) |", "|---|---|---|---|---|---|---|---|"]
for n in (2, 4):
    for arm in ("ours", "reference"):
        d = load(f"c4_n{n}_{arm}.json")
        if not d:
            continue
        lost = d.get("training_time_lost_per_snapshot_ms")
        if lost is None:
            lost = d["value"] + max(0.0, (d["step_ms_during_drain"] - d["train_step_ms"])) * d["config"]["steps_during_drain"]
        out.append(f"| {n} | {arm} | **{f(d['value'], 0)}** | {f(d['train_step_ms'])} | {f(d['step_ms_during_drain'])} | {f(d.get('overlap_pct'))} | {f(lost, 0)} | {f(d.get('sync_take_ms'), 0)} |")
out += ["", "overlap % = 1 − (extra time of the steps that ran while the snapshot drained) / (time from async_take returning to the snapshot being complete).  The reference does",
        "its whole D2H inside the blocking window and only writes files in the background, so its background phase disturbs training less — after stopping it 5–6x longer.", ""]
out += ["## C5 — 16 GB row-wise ShardedTensor saved at N, restored at N/2 (reshard-on-load)", "",
        "| N → N/2 | arm | save GB/s | reshard restore GB/s | restore ms | every element verified |", "|---|---|---|---|---|---|"]
for n in (2, 8):
    for arm in ("ours", "reference"):
        d = load(f"c5_n{n}_{arm}.json")
        if d:
            out.append(f"| {n} → {max(1, n // 2)} | {arm} | {f(d['save']['value'])} | **{f(d['value'])}** | {f(d['ms_per_step'], 0)} | {d.get('verified_every_element_on_restoring_ranks')} |")
open(os.path.join(ROOT, "profiles", "r02_bench_summary.md"), "w").write("\n".join(out) + "\n")

# ---- kernel cases -----------------------------------------------------------------------------------------------------
kc = ["# r02 — every kernel mode in isolation (CUDA events, `tools/kernel_cases.py`, 1 GiB-class members, N=1)", "",
      "`frac` = algorithmic bytes (read + written) / kernel time / 6565.8 GB/s (MEASURED_PEAKS.json).  Variants: LSU kernel bounded for 2 (128 regs) or 3 (80 regs) CTAs per SM;",
      "copy-engine rows threshold 128 B instead of 256 B (A/B columns: all LSU modes on one build, 32 KiB transpose tiles with scalar shared-memory accesses).",
      "Shipping configuration: strided tiles on the 2-CTA build, transpose tiles (16 KiB, conflict-free word layout) on the 6-CTA build, every other LSU mode on the 3-CTA build, rows ≥ 256 B.", ""]
files = sorted(glob.glob(os.path.join(RAW, "r02_kernel_cases_*.jsonl")))
cases = {}
for fn in files:
    tag = os.path.basename(fn)[len("r02_kernel_cases_"):-len(".jsonl")]
    for ln in open(fn):
        try:
            d = json.loads(ln)
        except Exception:
            continue
        cases.setdefault(d["case"], {})[tag] = d
tags = sorted({t for c in cases.values() for t in c})
kc += ["| case | " + " | ".join(f"{t}: ms / GB/s / frac" for t in tags) + " |", "|---|" + "---|" * len(tags)]
for name, per in cases.items():
    kc.append(f"| {name} | " + " | ".join((f"{per[t]['kernel_ms']:.4f} / {per[t]['gbs']:.0f} / **{per[t]['frac_of_measured_peak']:.3f}**" if t in per else "—") for t in tags) + " |")
open(os.path.join(ROOT, "profiles", "r02_kernel_cases.md"), "w").write("\n".join(kc) + "\n")
print("\n".join(out[:40]))
