"""The tensor-map TMA kernels against the kernels they replace, through the C-ABI stager seam: transposes (TMA tile kernel vs
the LSU tiled transpose, TSNAP_B200_TMA_TRANSPOSE=0) and short-run column shards (one TMA box of many runs vs one copy-engine
request per run, TSNAP_B200_TMA_ROWS=0).  CUDA-event kernel time per case, the staged image checked against `.contiguous()`.

    python tools/transpose_tune.py [--reps 3] [--cases a,b] > gpurun_out/transpose_cases.jsonl
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torchsnapshot_b200 import _native as N

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3); ap.add_argument("--cases", default="all"); ap.add_argument("--modes", default="1,0")
args = ap.parse_args()
dev = "cuda:0"
eng = N.Engine(device=0, io_threads=2, pinned_slot_bytes=32 << 20, pinned_slots=4)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0

base = torch.empty((1 << 30) // 4, dtype=torch.float32, device=dev).uniform_()  # 1 GiB
hb = base.view(torch.bfloat16)
db = base.view(torch.float64)


def cases():
    yield "fp32_8192x8192", base[: 8192 * 8192].view(8192, 8192).t(), True
    yield "fp32_8192x8200", base[: 8192 * 8200].view(8192, 8200).t(), True          # non-power-of-two pitch, edge tiles
    yield "fp32_1000x1004_edges", base[: 1000 * 1004].view(1000, 1004).t(), True    # every tile row/column has a clipped edge
    yield "fp32_16384x16384", base.view(16384, 16384).t(), False
    yield "bf16_16384x8192", hb[: 16384 * 8192].view(16384, 8192).t(), True
    yield "bf16_4096x14336", hb[: 4096 * 14336].view(4096, 14336).t(), True         # Llama-3-8B w2, transposed
    yield "bf16_3d_64x1024x1040", hb[: 64 * 1024 * 1040].view(64, 1024, 1040).transpose(1, 2), True
    yield "bf16_4d_permute", hb[: 8 * 12 * 520 * 264].view(8, 12, 520, 264).permute(0, 3, 1, 2), True
    yield "fp64_4096x6144", db[: 4096 * 6144].view(4096, 6144).t(), True
    n = (1 << 20) + 64
    yield "fp32_Nx64_tall", base[: n * 64].view(n, 64).t(), False
    yield "fp32_64xN_wide", base[: n * 64].view(64, n).t(), False
    # column shards: runs of 256 B .. 2 KiB out of rows twice as long (256 MiB of payload each)
    for run in (256, 512, 1024, 2048):
        rows = (512 << 20) // (2 * run)
        yield f"rows_{run}B_runs", base[: rows * 2 * run // 4].view(rows, 2 * run // 4)[:, run // 8 : run // 8 + run // 4], run == 512
    yield "rows_3d_272B_runs", base[: 50 * 1000 * 200].view(50, 1000, 200)[3:47, 5:990, 64:132], True  # 272 B runs, clipped boxes


for name, t, verify in cases():
    if args.cases != "all" and name not in args.cases.split(","):
        continue
    nbytes = t.numel() * t.element_size()
    want = t.contiguous().view(torch.uint8).reshape(-1).cpu() if verify else None
    for mode in args.modes.split(","):
        os.environ["TSNAP_B200_TMA_TRANSPOSE"] = mode
        os.environ["TSNAP_B200_TMA_ROWS"] = mode
        desc = [N.save_desc(t, 0)]
        best, ok = None, None
        for rep in range(args.reps + 1):
            sb = eng.stage(desc, nbytes + 16, stream=torch.cuda.current_stream().cuda_stream, keepalive=[t])
            mv = sb.wait(); st = sb.stats()
            if rep == 0 and want is not None:
                ok = bool(torch.equal(torch.frombuffer(mv, dtype=torch.uint8, count=nbytes), want))
            del mv
            sb.release()
            if rep and (best is None or st["kernel_ms"] < best):
                best = st["kernel_ms"]
        print(json.dumps({"case": name, "kernel": "tma" if mode != "0" else ("per-run" if name.startswith("rows") else "lsu"), "kernel_ms": round(best, 4), "gbs": round(2 * nbytes / 1e6 / best, 1),
                          "frac": round(2 * nbytes / 1e6 / best / peak, 3), "ok": ok}), flush=True)
eng.close()
