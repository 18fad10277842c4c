"""Run one sb200_time_phase on an scpnrh-shaped model (for ncu captures of a single kernel)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import sypha_b200 as sb
from sypha_b200 import _lib as L
from sypha_b200.instances import gen_scp
phase = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = L.load()
mdl = gen_scp(1000, 10000, 0.05, 1)
node = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, sb.SyphaEnvironment())
ws = sb.IpmWorkspace(); sb.initializeIpmWorkspace(ws); node.copyModelOnDevice(ws)
ms = C.c_double(); rc = lib.sb200_time_phase(ws.handle, phase, reps, C.byref(ms))
print(f"phase {phase}: rc {rc} {1e3*ms.value:.1f} us")
