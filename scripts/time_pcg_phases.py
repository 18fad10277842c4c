"""Device timing (CUDA events, sb200_time_phase) of the PCG building blocks on one synthetic instance."""
import argparse
import ctypes as C
import sys

sys.path.insert(0, ".")
import sypha_b200 as sb
from sypha_b200 import _lib as L
from sypha_b200.instances import gen_scp

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=50000)
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--density", type=float, default=0.001)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
lib = L.load()
mdl = gen_scp(a.m, a.n, a.density, 1)
env = sb.SyphaEnvironment(linearSolverStrategy="pcg")
node = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
ws = sb.IpmWorkspace()
sb.initializeIpmWorkspace(ws)
node.copyModelOnDevice(ws)
info = (C.c_longlong * 20)()
lib.sb200_model_info(ws.handle, info, 20)
print(f"m={mdl.m} n={mdl.n} nnz={mdl.nnz} blocked={info[12]} row-side blocks={info[13]} col-side blocks={info[14]} "
      f"chunks(16 B) row-side {info[15]} col-side {info[16]}")
names = {3: "A v (rhs SpMV)", 4: "A' v + recover + ratio test", 6: "one CG iteration", 7: "CG: q = D A'p", 8: "CG: Ap = A q (+ p.Ap)"}
for pid, nm in names.items():
    ms = C.c_double()
    rc = lib.sb200_time_phase(ws.handle, pid, a.reps, C.byref(ms))
    print(f"phase {pid} {nm:32s} rc={rc} {1e3*ms.value:9.1f} us")
sb.releaseIpmWorkspace(ws)
