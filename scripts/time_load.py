"""Host-side timing of sb200_load_model + solve on one scpnrh-shaped instance, repeated on one workspace."""
import sys, time
sys.path.insert(0, ".")
import sypha_b200 as sb
from sypha_b200.instances import gen_scp
mdl = gen_scp(1000, 10000, 0.05, 1)
env = sb.SyphaEnvironment()
ws = sb.IpmWorkspace(); sb.initializeIpmWorkspace(ws)
for rep in range(4):
    node = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    t0 = time.perf_counter(); node.copyModelOnDevice(ws); t1 = time.perf_counter()
    res = sb.SolverExecutionResult()
    sb.solver_sparse_mehrotra_run(node, sb.SolverExecutionConfig(maxIterations=100), res, ws)
    t2 = time.perf_counter()
    print(f"rep {rep}: load {1e3*(t1-t0):.1f} ms  solve {1e3*(t2-t1):.1f} ms  (start {res.msStart:.2f} setup {res.msSetup:.2f} loop {res.msLoop:.2f}) iters {res.iterations}")
