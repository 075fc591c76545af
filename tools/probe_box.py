"""One-off probe of the GPU box: host cores, memory, disks, pinned/pageable link bandwidth,
pwrite/pread throughput. Output goes to gpurun_out/probe_box.json. Not part of the product."""
import json, os, sys, time, threading, subprocess, shutil
import torch

out = {}
out["cpu_count"] = os.cpu_count()
try:
    out["sched_affinity"] = len(os.sched_getaffinity(0))
except Exception as e:
    out["sched_affinity"] = str(e)
import psutil
vm = psutil.virtual_memory()
out["mem_total_gb"] = vm.total / 2**30
out["mem_avail_gb"] = vm.available / 2**30
def sh(c):
    try:
        return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=60).stdout
    except Exception as e:
        return str(e)
out["df"] = sh("df -h /tmp /dev/shm /root/repo . 2>&1")
out["mounts"] = sh("mount | grep -E ' / | /tmp | /dev/shm |nvme|overlay' | head -20")
out["lsblk"] = sh("lsblk -d -o NAME,SIZE,ROTA,MODEL 2>&1 | head -20")
out["nvidia_smi"] = sh("nvidia-smi --query-gpu=index,name,memory.total,pcie.link.gen.current,pcie.link.width.current --format=csv")
out["topo"] = sh("nvidia-smi topo -m 2>&1 | head -30")
out["ulimit_l"] = sh("ulimit -l")
out["lscpu"] = sh("lscpu | grep -E 'Model name|Socket|NUMA|Thread|Core' ")

dev = torch.device("cuda:0")
torch.cuda.init()
out["gpu_name"] = torch.cuda.get_device_name(0)
out["gpu_count"] = torch.cuda.device_count()

def ev_time(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 1e3)
    return best

link = {}
for mb in (1, 16, 64, 256, 1024):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.uint8, pin_memory=True); t_pin = time.perf_counter() - t0
    h.fill_(1)
    d2h = ev_time(lambda: h.copy_(d, non_blocking=True))
    h2d = ev_time(lambda: d.copy_(h, non_blocking=True))
    # pageable, as the reference does it
    torch.cuda.synchronize(); t0 = time.perf_counter(); c = d.to("cpu"); t_pg = time.perf_counter() - t0
    t0 = time.perf_counter(); c = d.to("cpu"); t_pg2 = time.perf_counter() - t0
    pg = torch.empty(n, dtype=torch.uint8); pg.fill_(1)
    t0 = time.perf_counter(); pg.copy_(d); t_pg3 = time.perf_counter() - t0
    t0 = time.perf_counter(); d.copy_(pg); torch.cuda.synchronize(); t_h2d_pg = time.perf_counter() - t0
    link[mb] = dict(pin_alloc_s=t_pin, d2h_pinned_gbs=n / d2h / 1e9, h2d_pinned_gbs=n / h2d / 1e9,
                    d2h_pageable_fresh_gbs=n / t_pg / 1e9, d2h_pageable_fresh2_gbs=n / t_pg2 / 1e9,
                    d2h_pageable_warm_gbs=n / t_pg3 / 1e9, h2d_pageable_gbs=n / t_h2d_pg / 1e9)
    del d, h, c, pg
out["link"] = link

# two concurrent D2H streams + bidirectional
n = 512 << 20
d1 = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
def two_d2h():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
def bidir():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
out["d2h_two_streams_gbs"] = 2 * n / ev_time(two_d2h) / 1e9
out["bidir_gbs_total"] = 2 * n / ev_time(bidir) / 1e9

# file write / read throughput from a pinned buffer, T threads, pwrite of 8 MiB blocks
import numpy as np
hb = h1.numpy()
def file_bw(root, threads, total=4 << 30, block=8 << 20, direct=False, sync=False):
    os.makedirs(root, exist_ok=True)
    per = total // threads
    paths = [os.path.join(root, f"probe_{i}.bin") for i in range(threads)]
    def w(i):
        flags = os.O_WRONLY | os.O_CREAT | os.O_TRUNC
        if direct: flags |= os.O_DIRECT
        fd = os.open(paths[i], flags, 0o644)
        off = 0
        mv = memoryview(hb)
        while off < per:
            k = min(block, per - off)
            src = (off % (n - block))
            src -= src % 4096
            os.pwrite(fd, mv[src:src + k], off); off += k
        if sync: os.fsync(fd)
        os.close(fd)
    ts = [threading.Thread(target=w, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]; tw = time.perf_counter() - t0
    def r(i):
        fd = os.open(paths[i], os.O_RDONLY)
        off = 0
        buf = memoryview(h2.numpy())
        while off < per:
            k = min(block, per - off)
            dst = off % (n - block)
            os.preadv(fd, [buf[dst:dst + k]], off); off += k
        os.close(fd)
    ts = [threading.Thread(target=r, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]; tr = time.perf_counter() - t0
    for p in paths: os.unlink(p)
    return dict(write_gbs=total / tw / 1e9, read_cached_gbs=total / tr / 1e9)
fs = {}
for root in ("/tmp/tsnap_probe", "/dev/shm/tsnap_probe"):
    for th in (1, 4, 8, 16):
        try:
            fs[f"{root}:t{th}"] = file_bw(root, th)
        except Exception as e:
            fs[f"{root}:t{th}"] = str(e)
    for th in (4, 16):
        try:
            fs[f"{root}:t{th}:fsync"] = file_bw(root, th, sync=True)
        except Exception as e:
            fs[f"{root}:t{th}:fsync"] = str(e)
    try:
        fs[f"{root}:t8:odirect"] = file_bw(root, 8, direct=True)
    except Exception as e:
        fs[f"{root}:t8:odirect"] = str(e)
    shutil.rmtree(root, ignore_errors=True)
out["fs"] = fs
# host memcpy bandwidth (numpy) single thread
a = np.empty(1 << 30, dtype=np.uint8); b = np.ones(1 << 30, dtype=np.uint8)
t0 = time.perf_counter(); a[:] = b; t1 = time.perf_counter() - t0
t0 = time.perf_counter(); a[:] = b; t2 = time.perf_counter() - t0
out["host_memcpy_gbs"] = [(1 << 30) / t1 / 1e9, (1 << 30) / t2 / 1e9]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_box.json", "w"), indent=1)
print(json.dumps(out, indent=1))
