"""Profiling driver: factor + solve one dense SPD matrix through the L0 entry points."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sypha_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
ld = (n + 63) // 64 * 64
r = np.random.default_rng(0)
B = r.normal(size=(n, n + 8))
M = np.eye(ld)
M[:n, :n] = B @ B.T + 1e-3 * n * np.eye(n)
Md = torch.from_numpy(M).cuda()
info = torch.zeros(1, dtype=torch.int32, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
b = torch.zeros(ld, dtype=torch.float64, device="cuda")
b[:n] = 1.0
for it in range(reps):
    A = Md.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lib.sb200_k_potrf(n, C.c_void_p(A.data_ptr()), ld, C.c_void_p(info.data_ptr()), st)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    x = b.clone()
    lib.sb200_k_potrs(n, C.c_void_p(A.data_ptr()), ld, C.c_void_p(x.data_ptr()), st)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"n={n} potrf {1e6*(t1-t0):.1f} us  potrs {1e6*(t2-t1):.1f} us  info={int(info.item())}")
ref = np.linalg.solve(M[:n, :n], np.ones(n))
print("max rel err", float(np.max(np.abs(x.cpu().numpy()[:n] - ref)) / np.abs(ref).max()))
