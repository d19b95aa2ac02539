import ctypes, sys
rt = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_int()
for name, attr in (("maxPersistingL2", 108), ("maxAccessPolicyWindow", 109), ("l2CacheSize", 38)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0); print(name, v.value)
