"""Did the iteration graph get captured (model_info[10] = kernels per captured iteration, 0 = stream launches)?"""
import ctypes as C, sys
sys.path.insert(0, ".")
import sypha_b200 as sb
from sypha_b200 import _lib as L
from sypha_b200.instances import gen_scp
lib = L.load()
mdl = gen_scp(1000, 10000, 0.05, 1)
env = sb.SyphaEnvironment()
node = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
ws = sb.IpmWorkspace(); sb.initializeIpmWorkspace(ws)
res = sb.SolverExecutionResult()
sb.solver_sparse_mehrotra_run(node, sb.SolverExecutionConfig(maxIterations=100), res, ws)
info = (C.c_longlong * 20)(); lib.sb200_model_info(ws.handle, info, 20)
print("iterations", res.iterations, "primal", res.primalObj, "graph kernels per iteration", info[10], "last_error:", lib.sb200_last_error(ws.handle).decode()[:120])
