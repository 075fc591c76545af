"""Times the pack kernel alone (CUDA events inside the engine) for one geometry; run once per TSNAP_B200_BULK_CFG."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torchsnapshot_b200 import _native as N
dev = torch.device("cuda:0")
eng = N.get_engine(0)
ts = [torch.empty(1 << 28, dtype=torch.bfloat16, device=dev).normal_() for _ in range(16)]  # 16 x 512 MiB = 8 GiB
descs, off = [], 0
for t in ts:
    descs.append(N.save_desc(t, off)); off += t.numel() * 2
ms = []
for i in range(6):
    sb = eng.stage(descs, off, stream=torch.cuda.current_stream().cuda_stream, keepalive=ts); sb.wait(); st = sb.stats(); sb.release()
    if i >= 2: ms.append(st["kernel_bulk_ms"])
best = min(ms)
print(json.dumps({"cfg": os.environ.get("TSNAP_B200_BULK_CFG", "0"), "ms": [round(x, 3) for x in ms], "best_gbs": round(2 * off / 1e9 / (best / 1e3), 1), "frac_of_6565.8": round(2 * off / 1e9 / (best / 1e3) / 6565.8, 4)}))
