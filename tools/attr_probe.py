"""Device attributes that decide whether a kernel can read cudaHostRegister'ed memory through the host pointer."""
import ctypes
import glob
import os

import torch

torch.cuda.init()
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
    glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = ctypes.CDLL(cands[0])
for name, k in [("CanMapHostMemory", 19), ("UnifiedAddressing", 41), ("CanUseHostPointerForRegisteredMem", 91),
                ("HostRegisterSupported", 99), ("PageableMemoryAccess", 88), ("HostNativeAtomicSupported", 86)]:
    v = ctypes.c_int(-1)
    rc = rt.cudaDeviceGetAttribute(ctypes.byref(v), k, 0)
    print(name, v.value, "rc", rc)
