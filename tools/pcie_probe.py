"""PCIe copy-bandwidth probe (diagnostic, not part of the product): pinned-alloc vs cudaHostRegister'ed
pageable memory, H2D and D2H, one copy vs several concurrent ones."""
import ctypes
import subprocess
import time

import numpy as np
import torch

rt = torch.cuda.cudart()
dev = torch.device("cuda", 0)
N = 1 << 30  # bytes


def t_copy(dst, src, stream=None, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return src.numel() * src.element_size() / best / 1e9


d = torch.empty(N, dtype=torch.uint8, device=dev)
hp = torch.empty(N, dtype=torch.uint8).pin_memory()
hp.fill_(1)
print("pinned-alloc  H2D %.1f GB/s  D2H %.1f GB/s" % (t_copy(d, hp), t_copy(hp, d)))

a = np.ones(N, dtype=np.uint8)
hr = torch.from_numpy(a)
r = rt.cudaHostRegister(a.ctypes.data, N, 0)
print("cudaHostRegister rc", r)
print("registered    H2D %.1f GB/s  D2H %.1f GB/s" % (t_copy(d, hr), t_copy(hr, d)))
# registered std::vector-like (malloc'ed, touched by one thread)
t0 = time.perf_counter()
b = np.ones(N, dtype=np.uint8)
t1 = time.perf_counter()
r = rt.cudaHostRegister(b.ctypes.data, N, 0)
t2 = time.perf_counter()
print("alloc+touch 1 GiB %.0f ms, register %.0f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
# full duplex
d2 = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s1):
    d.copy_(hp, non_blocking=True)
with torch.cuda.stream(s2):
    hr.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("duplex: 1 GiB each way in %.1f ms -> %.1f GB/s per direction" % (dt * 1e3, N / dt / 1e9))
for cmd in ("nvidia-smi topo -m", "lscpu | grep -i -E 'numa|model name|socket|^cpu\\(s\\)'", "free -g | head -2"):
    print(subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout)
# large registrations: does cudaHostRegister cope with > 1 GiB in one call, and how fast is the copy?
for gib in (1.6, 3.2):
    n = int(gib * (1 << 30))
    big = np.ones(n, dtype=np.uint8)
    r = rt.cudaHostRegister(big.ctypes.data, n, 0)
    dd = torch.empty(n, dtype=torch.uint8, device=dev)
    print("register %.1f GiB rc=%s  H2D %.1f GB/s" % (gib, r, t_copy(dd, torch.from_numpy(big))))
    rt.cudaHostUnregister(big.ctypes.data)
    del dd, big
