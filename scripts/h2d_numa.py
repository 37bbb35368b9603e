"""Host-to-device bandwidth of pinned buffers allocated under the default CPU affinity vs the GPU-local one."""
import os, time
import torch
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
ncpu = os.cpu_count()
words = (ncpu + 63) // 64
try:
    aff = pynvml.nvmlDeviceGetCpuAffinity(h, words)
    local = [i for i in range(ncpu) if (aff[i // 64] >> (i % 64)) & 1]
except Exception as e:
    local = []
    print("no affinity info:", e)
print("cpus", ncpu, "default affinity", len(os.sched_getaffinity(0)), "gpu-local cpus", len(local), local[:4], "...")
dev = torch.device("cuda:0")
dst = torch.empty(64 << 20, dtype=torch.float32, device=dev)

def bw(tag):
    src = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
    src.fill_(1.0)
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s: %.1f GB/s" % (tag, 10 * src.numel() * 4 / dt / 1e9), flush=True)

bw("default affinity")
if local:
    allc = os.sched_getaffinity(0)
    os.sched_setaffinity(0, set(local) & allc or allc)
    bw("gpu-local affinity")
    far = sorted(allc - set(local))
    if far:
        os.sched_setaffinity(0, set(far))
        bw("remote cpus")
    os.sched_setaffinity(0, allc)
