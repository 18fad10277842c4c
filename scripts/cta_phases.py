#!/usr/bin/env python
"""Per-phase time inside the one-CTA solver (sb200_cta.cu), one LP alone and K of them in flight.
   python scripts/cta_phases.py [K]"""
import ctypes as C
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import sypha_b200 as sb  # noqa: E402
from sypha_b200 import _lib as L, bnb, solver as S  # noqa: E402
from sypha_b200.instances import load_npz  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
lib = L.load()
names = ["assembly", "factorisation", "solves", "A v", "A' v", "vector", "start point", "whole LP"]
for inst in ("scpnre1", "scpnrg1", "scpnrh1"):
    mdl = load_npz(REPO / "tests" / "golden" / f"{inst}.npz")
    if inst == "scpnre1":
        mdl, _ = bnb.reduce_by_incumbent(mdl, 38.0)
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=100)
    base = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    for k in (1, K):
        wss = [S.workspace_for_nodes(base, 8) for _ in range(k)]
        for w in wss:
            S.set_solver_form(w, "throughput")
        S.solve_batch_nodes(base, [()] * k, cfg, wss, fetch_solutions=False)
        t0 = time.perf_counter()
        res = S.solve_batch_nodes(base, [()] * k, cfg, wss, fetch_solutions=False)
        wall = time.perf_counter() - t0
        ms = C.c_double()
        ph = []
        for i in range(8):
            lib.sb200_time_phase(wss[k // 2].handle, 100 + i, 1, C.byref(ms))
            ph.append(ms.value)
        sub = []
        for i in range(3):
            lib.sb200_time_phase(wss[k // 2].handle, 108 + i, 1, C.byref(ms))
            sub.append(ms.value)
        it = res[0].iterations
        print(f"{inst} m={mdl.m} n={mdl.n} nnz={mdl.nnz}: {k} LPs in flight, {it} iterations, wall {1e3 * wall:.1f} ms "
              f"({k * it / wall:.0f} iter/s aggregate); per iteration (us): "
              + ", ".join(f"{nm} {1e3 * v / it:.0f}" for nm, v in zip(names[:6], ph[:6]))
              + f"; start point {1e3 * ph[6]:.0f} us, whole LP {ph[7]:.2f} ms; factorisation split per factorisation (us): "
              + ", ".join(f"{nm} {1e3 * v / (it + 1):.0f}" for nm, v in zip(("accumulate", "diagonal tiles", "epilogues"), sub)))
        for w in wss:
            sb.releaseIpmWorkspace(w)
