"""ncu report -> committed evidence: raw per-launch CSV (selected metrics) + a markdown table.

    python tools/ncu_summarize.py gpurun_out/r02_kernels.ncu-rep profiles/r02_ncu_kernels [case names...]
"""
import csv, io, json, os, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
labels = sys.argv[3:]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
with open(out + "_raw.csv", "w") as f:
    f.write(raw)
ix = {n: i for i, n in enumerate(hdr)}
def g(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return None
unit = lambda k: units[ix[k]] if k in ix else ""
lines = ["| # | kernel | case | grid x block | regs | time (us) | DRAM read + write (GB) | DRAM GB/s | L1 hit % | L2 hit % |", "|---|---|---|---|---|---|---|---|---|---|"]
for i, r in enumerate(data):
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("tsnap::", "")
    t = g(r, "gpu__time_duration.sum"); tu = unit("gpu__time_duration.sum")
    us = t / 1e3 if tu in ("ns", "nsecond") else (t if tu in ("us", "usecond") else t * 1e3)
    def gb(k):
        v = g(r, k); u = unit(k)
        return v / {"byte": 1e9, "Kbyte": 1e6, "Mbyte": 1e3, "Gbyte": 1.0}.get(u, 1e9)
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    per = max(1, len(data) // len(labels)) if labels else 1
    case = labels[i // per] if labels and i // per < len(labels) else ""
    lines.append(f"| {i} | `{name}` | {case} | {int(g(r, 'launch__grid_size'))} x {int(g(r, 'launch__block_size'))} | {int(g(r, 'launch__registers_per_thread'))} | {us:.1f} | {rd:.3f} + {wr:.3f} | {(rd + wr) / (us / 1e6):.0f} | {g(r, 'l1tex__t_sector_hit_rate.pct'):.1f} | {g(r, 'lts__t_sector_hit_rate.pct'):.1f} |")
with open(out + "_table.md", "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
