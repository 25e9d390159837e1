import ctypes, os, sys
sys.path.insert(0, "/root/repo")
import torch
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
lim = int(os.environ.get("PERSIST_MB", "0"))
if lim:
    v = ctypes.c_int()
    rt.cudaDeviceGetAttribute(ctypes.byref(v), 108, 0)  # cudaDevAttrMaxPersistingL2CacheSize
    rc = rt.cudaDeviceSetLimit(6, ctypes.c_size_t(min(lim << 20, v.value)))  # cudaLimitPersistingL2CacheSize
    got = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(got), 6)
    print(f"max persisting L2 {v.value>>20} MB, set rc={rc}, now {got.value>>20} MB")
exec(open("/root/repo/scripts/lookup_ablate.py").read())
