"""Exploration driver for the big synthetic instances: load one model, run a few IPM iterations with
the chosen strategy, print timings (used to size bench.py's synth50k workload)."""
import argparse
import sys
import time

sys.path.insert(0, ".")
import sypha_b200 as sb
from sypha_b200.instances import gen_scp

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=50000)
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--density", type=float, default=0.001)
ap.add_argument("--strategy", default="pcg")
ap.add_argument("--max-iter", type=int, default=3)
ap.add_argument("--cg-max-iter", type=int, default=50000)
ap.add_argument("--cg-tol", type=float, default=1e-8)
a = ap.parse_args()
t0 = time.time()
mdl = gen_scp(a.m, a.n, a.density, 1)
print(f"generated {mdl.m}x{mdl.n} nnz={mdl.nnz} in {time.time()-t0:.1f}s", flush=True)
env = sb.SyphaEnvironment(linearSolverStrategy=a.strategy, krylovMaxCgIter=a.cg_max_iter, krylovCgTolInitial=a.cg_tol,
                          krylovCgTolFinal=a.cg_tol, krylovCgTolDecayRate=1.0)
node = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
ws = sb.IpmWorkspace()
sb.initializeIpmWorkspace(ws)
t0 = time.time()
node.copyModelOnDevice(ws)
print(f"load_model {time.time()-t0:.2f}s", flush=True)
for it in (a.max_iter,):
    res = sb.SolverExecutionResult()
    t0 = time.time()
    st = sb.solver_sparse_mehrotra_run(node, sb.SolverExecutionConfig(maxIterations=it), res, ws)
    dt = time.time() - t0
    print(f"status {st} reason {res.terminationReason} iters {res.iterations} primal {res.primalObj:.6f} dual {res.dualObj:.6f} "
          f"mu {res.mu:.3e} wall {dt:.2f}s start {res.msStart:.1f}ms loop {res.msLoop:.1f}ms cg_iters {res.cgIterations} "
          f"kernels {res.kernelsLaunched}", flush=True)
    if res.cgIterations:
        print(f"  ~{1e3*(res.msStart+res.msLoop)/res.cgIterations:.1f} us per CG iteration", flush=True)
