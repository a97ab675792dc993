"""Where the start-up time of a context goes: CUDA context creation vs pf_create itself."""
import sys, time
sys.path.insert(0, ".")
from panfeed_b200 import capi
import ctypes
which = sys.argv[1] if len(sys.argv) > 1 else "pf_first"
if which == "torch_first":
    t = time.perf_counter(); import torch; torch.zeros(1, device="cuda"); torch.cuda.synchronize()
    print(f"torch CUDA init + first alloc: {time.perf_counter() - t:.2f} s")
for i in range(3):
    t = time.perf_counter(); ctx = capi.Context(31, 500, maf=0.01); dt = time.perf_counter() - t
    print(f"pf_create #{i}: {dt:.3f} s")
    t = time.perf_counter(); ctx.close(); print(f"  pf_destroy: {time.perf_counter() - t:.3f} s")
