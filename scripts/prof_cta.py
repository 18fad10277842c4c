#!/usr/bin/env python
"""The launches ncu profiles for the throughput form: scpnrh1 solved whole by one thread block (k_ipm_cta) a few times, the
node-heuristics kernel behind it, and for comparison the same LP in the latency form.
   python scripts/prof_cta.py [instance] [reps]"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import sypha_b200 as sb  # noqa: E402
from sypha_b200 import solver as S  # noqa: E402
from sypha_b200.instances import load_npz  # noqa: E402

inst = sys.argv[1] if len(sys.argv) > 1 else "scpnrh1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mdl = load_npz(REPO / "tests" / "golden" / f"{inst}.npz")
env = sb.SyphaEnvironment()
cfg = sb.SolverExecutionConfig(maxIterations=100)
base = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
ws = S.workspace_for_nodes(base, 8)
for form in ("throughput", "latency"):
    S.set_solver_form(ws, form)
    for _ in range(reps):
        r = S.solve_batch_nodes(base, [((3, 1), (40, 0))], cfg, [ws], fetch_solutions=False)[0]
        h = S.node_heuristics([ws])[0]
    print(form, r.iterations, r.primalObj, r.dualObj, int(r.kernelsLaunched), h.coverObj, h.repairSteps)
sb.releaseIpmWorkspace(ws)
