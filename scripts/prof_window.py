#!/usr/bin/env python
"""The launch ncu profiles for the throughput headline: windows of K LPs of scpnrh1-5 (one thread block each, ONE launch
of k_ipm_cta per window).   python scripts/prof_window.py [K] [windows]"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import sypha_b200 as sb  # noqa: E402
from sypha_b200 import solver as S  # noqa: E402
from sypha_b200.instances import load_npz  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 148
windows = int(sys.argv[2]) if len(sys.argv) > 2 else 2
models = [load_npz(REPO / "tests" / "golden" / f"scpnrh{i}.npz") for i in range(1, 6)]
env = sb.SyphaEnvironment()
cfg = sb.SolverExecutionConfig(maxIterations=100)
wss, nodes = [], []
for i in range(K):
    mdl = models[i % 5]
    w = sb.IpmWorkspace()
    sb.initializeIpmWorkspace(w)
    nd = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    nd.copyModelOnDevice(w)
    S.set_solver_form(w, "throughput")
    wss.append(w)
    nodes.append(nd)
for _ in range(windows):
    rs = S.solve_batch(nodes, cfg, wss)
    ms, k = S.last_window(wss[0])
    its = sum(r.iterations for r in rs)
    print(f"window of {k} LPs: {ms:.2f} ms, {its} iterations ({its / ms * 1e3:.0f} iter/s), launches {sum(int(r.kernelsLaunched) for r in rs)}")
for w in wss:
    sb.releaseIpmWorkspace(w)
