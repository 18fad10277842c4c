#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE's own code.

Runs only in the build container (needs /root/reference).  For every instance it
  1. reads the OR-Library file with the reference's reader (python/model_importer.py),
  2. runs the reference's Mehrotra solver, unmodified, with the C++ solver's eta = 0.95
     (src/sypha_environment_defaults.h:15) for exactly ``k`` iterations, k = the number of
     iterations the C++ loop test ``mu > 1e-4`` (src/sypha_solver.cpp:496) allows - the Python
     prototype's own loop test is hard-coded to ``mu > 1e-10`` (python/interior_point.py:106,442),
     so it is truncated through its ``k_max`` argument instead of being patched,
  3. stores the instance (compact CSR of A0), the reference's final (x, y, s) and objectives, and
     the known LP optimum from python/sypha_unit_tests.py:21-77 /
     benchmark/results/benchmark_results_with_ip.csv.

  small instances : interior_point.mehrotra_linopt_dense
  large instances : interior_point.mehrotra_linopt_sparse   (slow: minutes each; ``--large``)

Usage:  python tests/golden/make_golden.py [--large] [names...]
"""
import argparse
import csv
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF / "python"))

SMALL = ["demo00", "scp_demo06", "scp_demo_tiny03", "scp41", "scp42", "scp46", "scp48", "scp49",
         "scp410", "scp51", "scpclr10", "scpcyc06", "scpa1", "scpb1"]
LARGE = ["scpnre1", "scpnrf1", "scpnrg1", "scpnrh1", "scpclr13"]
# the rest of the families BASELINE.json configs[1] / configs[4] name (bench.py's default workloads)
LARGE += [f"scpnr{f}{i}" for f in "ehg" for i in range(2, 6)]


def known_lp_optima():
    """name -> LP optimum from the reference's two known-answer tables."""
    out = {}
    with open(REF / "benchmark/results/benchmark_results_with_ip.csv") as fh:
        for row in csv.DictReader(fh):
            if row["lp_status"] == "OPTIMAL":
                out[row["instance"].replace(".txt", "")] = float(row["lp_objective"])
    # python/sypha_unit_tests.py:21-77 (20-digit tables) take precedence
    ns = {}
    src = (REF / "python/sypha_unit_tests.py").read_text()
    head = src.split("def exec_sypha")[0]
    exec(compile(head.split("from argparse import ArgumentParser")[1], "gold", "exec"), ns)
    for fam, key in (("scp4", "SCP4"), ("scp5", "SCP5"), ("scpnre", "SCPNRE"), ("scpnrf", "SCPNRF"),
                     ("scpnrg", "SCPNRG"), ("scpnrh", "SCPNRH")):
        for nm, val in zip(ns[key + "_STRINGS"], ns[key + "_SOLUTIONS"]):
            out[nm] = float(val)
    return out


def known_ip_optima():
    """name -> integer optimum (SCIP, status OPTIMAL) held by the reference:
    benchmark/results/benchmark_results_with_ip.csv, columns ip_status / ip_objective."""
    out = {}
    with open(REF / "benchmark/results/benchmark_results_with_ip.csv") as fh:
        for row in csv.DictReader(fh):
            if row.get("ip_status") == "OPTIMAL":
                out[row["instance"].replace(".txt", "")] = float(row["ip_objective"])
    return out


def run_one(name, large):
    import interior_point as ip
    import model_importer as mi
    from oracle import mehrotra as mo
    from oracle import scp_io

    path = REF / "data" / f"{name}.txt"
    inst = scp_io.load_scp(path, name)
    # iteration count under the C++ stop test, from the oracle (cross-checked below)
    t0 = time.time()
    orc = mo.solve_instance(inst, mo.Params(max_iter=100), "ne")
    t_or = time.time() - t0
    k = orc.iterations

    t0 = time.time()
    if not large:
        mat, rhs, obj = mi.sc_dense_model_reader(path)
        A, b, c = mi.sc_dense_to_standard_form(mat, rhs, obj)
        assert np.array_equal(A, inst.dense())
        x, y, s, it = ip.mehrotra_linopt_dense(A, b, c.astype(float), eta=0.95, k_max=k)
        kind = "interior_point.mehrotra_linopt_dense"
    else:
        mat, rhs, obj = mi.sc_sparse_model_reader(path)
        A, b, c = mi.sc_sparse_to_standard_form(mat, rhs, obj)
        A = A.tocsr()
        assert (abs(A - inst.csr())).nnz == 0
        x, y, s, it = ip.mehrotra_linopt_sparse(A, b, c.astype(float), eta=0.95, k_max=k)
        kind = "interior_point.mehrotra_linopt_sparse"
    t_ref = time.time() - t0
    assert it == k
    n0 = inst.n_orig
    mu = float(x @ s) / inst.n
    assert mu <= 1e-4, (name, mu)           # the C++ loop would indeed have stopped here ...
    primal = float(x[:n0] @ inst.c[:n0])
    dual = float(y @ inst.b)

    rows_offs = (inst.offs - np.arange(inst.m + 1)).astype(np.int32)     # strip the surplus entries
    mask = np.ones(inst.nnz, dtype=bool)
    mask[inst.offs[1:] - 1] = False
    cols = inst.inds[mask]
    cols = cols.astype(np.uint16) if n0 <= 65535 else cols.astype(np.int32)
    out = REPO / "tests/golden" / f"{name}.npz"
    np.savez_compressed(
        out, m=inst.m, n_orig=n0, costs=inst.c[:n0].astype(np.float32), row_offs=rows_offs,
        col_inds=cols, ref_kind=kind, ref_iters=k, ref_x=x, ref_y=y, ref_s=s,
        ref_primal=primal, ref_dual=dual, ref_mu=mu, ref_seconds=t_ref,
        oracle_primal=orc.primal, oracle_dual=orc.dual,
        oracle_mu_trace=np.array([t["mu"] for t in orc.trace]),
        oracle_alpha_p=np.array([t["alpha_p"] for t in orc.trace]),
        oracle_alpha_d=np.array([t["alpha_d"] for t in orc.trace]),
        lp_gold=known_lp_optima().get(name, np.nan))
    print(f"{name}: k={k} ref primal={primal:.9f} dual={dual:.9f} mu={mu:.3e} "
          f"oracle primal={orc.primal:.9f} dual={orc.dual:.9f} "
          f"[ref {t_ref:.1f}s oracle {t_or:.1f}s] -> {out.name} ({out.stat().st_size/1024:.0f} KB)",
          flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--large", action="store_true")
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    names = a.names or (LARGE if a.large else SMALL)
    for nm in names:
        run_one(nm, nm in LARGE)
    with open(REPO / "tests/golden/lp_optima.json", "w") as fh:
        json.dump(known_lp_optima(), fh, indent=0, sort_keys=True)
    with open(REPO / "tests/golden/ip_optima.json", "w") as fh:
        json.dump(known_ip_optima(), fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
