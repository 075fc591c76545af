"""Timeline of an engine job (``TSNAP_B200_ENGINE_FLAGS`` bit 8 / ``ENGINE_TRACE``): the overlap evidence of the
pack ‖ D2H ‖ pwrite pipeline (and pread ‖ H2D ‖ scatter on restore) without an external timeline profiler.

The reference's counterpart is the progress table of ``_WriteReporter`` (T:scheduler.py:98-177): staged / written
byte counts per tick.  Here every chunk leaves a record (``tsnap_job_get_trace``); this module turns them into
busy intervals per stage, pairwise overlap, and a Chrome-trace JSON (chrome://tracing, Perfetto)."""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Tuple

Interval = Tuple[float, float]


def _merge(iv: Iterable[Interval]) -> List[Interval]:
    out: List[Interval] = []
    for a, b in sorted(iv):
        if b <= a:
            continue
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], b))
        else:
            out.append((a, b))
    return out


def _total(iv: List[Interval]) -> float:
    return sum(b - a for a, b in iv)


def _intersect(x: List[Interval], y: List[Interval]) -> float:
    i = j = 0
    t = 0.0
    while i < len(x) and j < len(y):
        a, b = max(x[i][0], y[j][0]), min(x[i][1], y[j][1])
        if b > a:
            t += b - a
        if x[i][1] < y[j][1]:
            i += 1
        else:
            j += 1
    return t


def summarize(trace: List[dict]) -> Dict[str, object]:
    """Busy time (union of intervals) per stage, the span each stage covers, pairwise overlap and per-chunk figures."""
    by: Dict[str, List[dict]] = {}
    for r in trace:
        by.setdefault(r["kind"], []).append(r)
    out: Dict[str, object] = {}
    merged: Dict[str, List[Interval]] = {}
    for kind, recs in by.items():
        iv = _merge((r["t0_ms"], r["t1_ms"]) for r in recs)
        merged[kind] = iv
        nbytes = sum(r["bytes"] for r in recs)
        span = (min(r["t0_ms"] for r in recs), max(r["t1_ms"] for r in recs))
        busy = _total(iv)
        lanes = len({r["lane"] for r in recs})
        work = sum(r["t1_ms"] - r["t0_ms"] for r in recs)
        out[kind] = {
            "records": len(recs),
            "bytes": nbytes,
            "span_ms": [round(span[0], 3), round(span[1], 3)],
            "busy_ms": round(busy, 3),
            "lanes": lanes,
            "lane_ms_total": round(work, 3),
            "gbps_over_span": round(nbytes / 1e6 / max(span[1] - span[0], 1e-9), 2) if nbytes else None,
            "mean_ms_per_record": round(work / len(recs), 3),
        }
    pairs = (("kernel", "d2h"), ("d2h", "pwrite"), ("kernel", "pwrite"), ("pread", "h2d"), ("h2d", "kernel"))
    ov = {}
    for a, b in pairs:
        if a in merged and b in merged:
            both = _intersect(merged[a], merged[b])
            ov[f"{a}&{b}"] = {
                "both_busy_ms": round(both, 3),
                f"frac_of_{a}": round(both / max(_total(merged[a]), 1e-9), 3),
                f"frac_of_{b}": round(both / max(_total(merged[b]), 1e-9), 3),
            }
    out["overlap"] = ov
    if trace:
        out["job_span_ms"] = round(max(r["t1_ms"] for r in trace) - min(r["t0_ms"] for r in trace), 3)
    return out


def to_chrome_trace(trace: List[dict], label: str = "tsnap_b200 job") -> str:
    """Chrome trace-event JSON: one row per stage (I/O records get one row per worker)."""
    tid_of = {"plan": 1, "kernel": 2, "d2h": 3, "h2d": 3, "slot_wait": 4, "open": 5}
    ev = [{"name": "process_name", "ph": "M", "pid": 1, "args": {"name": label}}]
    names = {1: "plan (drain thread)", 2: "kernels (s_kernel)", 3: "link copies (s_copy)", 4: "slot wait (drain thread)", 5: "file open"}
    for r in trace:
        tid = tid_of.get(r["kind"])
        if tid is None:
            tid = 100 + r["lane"]
            names[tid] = f"io worker {r['lane']}"
        ev.append({"name": r["kind"], "ph": "X", "pid": 1, "tid": tid, "ts": r["t0_ms"] * 1e3, "dur": max(0.0, (r["t1_ms"] - r["t0_ms"]) * 1e3),
                   "args": {"bytes": r["bytes"], "file": r["file"]}})
    for tid, n in names.items():
        ev.append({"name": "thread_name", "ph": "M", "pid": 1, "tid": tid, "args": {"name": n}})
    return json.dumps({"traceEvents": ev, "displayTimeUnit": "ms"})
