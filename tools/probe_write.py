"""Characterise the host write path: buffered pwrite scaling with/without concurrent D2H DMA traffic."""
import json, os, sys, threading, time, shutil
import torch
dev = torch.device("cuda:0")
SLOT = 32 << 20
NSLOT = 32
ring = torch.empty(NSLOT * SLOT, dtype=torch.uint8, pin_memory=True); ring.fill_(7)
mv = memoryview(ring.numpy())
src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
stop = False
def dma_loop():
    s = torch.cuda.Stream()
    dst = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    n = 0
    with torch.cuda.stream(s):
        while not stop:
            dst.copy_(src, non_blocking=True); s.synchronize(); n += 1
    dma_loop.bytes = n << 30
def run(root, threads, total, nfiles, dma):
    global stop
    os.makedirs(root, exist_ok=True)
    chunks = total // SLOT
    fds = [os.open(os.path.join(root, f"f{i}"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644) for i in range(nfiles)]
    per_file = chunks // nfiles
    work = [(c % nfiles, c // nfiles) for c in range(per_file * nfiles)]  # interleaved across files
    idx = [0]; lock = threading.Lock()
    def w(tid):
        while True:
            with lock:
                i = idx[0]; idx[0] += 1
            if i >= len(work): return
            f, c = work[i]
            off = (i % NSLOT) * SLOT
            os.pwrite(fds[f], mv[off:off + SLOT], c * SLOT)
    stop = False
    t_dma = None
    if dma:
        t_dma = threading.Thread(target=dma_loop); t_dma.start(); time.sleep(0.2)
    ts = [threading.Thread(target=w, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]; dt = time.perf_counter() - t0
    stop = True
    if t_dma: t_dma.join()
    for fd in fds: os.close(fd)
    shutil.rmtree(root, ignore_errors=True)
    return round(len(work) * SLOT / dt / 1e9, 1)
out = {}
for root in ("/tmp/wprobe", "/dev/shm/wprobe"):
    for threads in (8, 16, 32, 64):
        for nfiles in (threads, 132):
            for dma in (False, True):
                k = f"{root} t{threads} files{nfiles} dma{int(dma)}"
                out[k] = run(root, threads, 16 << 30, nfiles, dma)
                print(k, out[k], flush=True)
json.dump(out, open("gpurun_out/probe_write.json", "w"), indent=1)
