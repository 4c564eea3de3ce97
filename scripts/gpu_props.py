#!/usr/bin/env python
"""Device attributes the design leans on (L2 size, persisting set-aside, access-policy window, shared memory)."""
import ctypes
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
def attr(i):
    v = ctypes.c_int()
    rc = rt.cudaDeviceGetAttribute(ctypes.byref(v), i, 0)
    return v.value if rc == 0 else None
names = {38: "l2CacheSize", 108: "maxPersistingL2CacheSize", 109: "maxAccessPolicyWindowSize", 81: "maxSharedMemoryPerMultiprocessor",
         97: "maxSharedMemoryPerBlockOptin", 16: "multiProcessorCount", 39: "maxThreadsPerMultiProcessor", 82: "maxRegistersPerMultiprocessor"}
print({n: attr(i) for i, n in names.items()})
