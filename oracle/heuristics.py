"""ctypes face of oracle/heur_oracle.c - the CPU restatement of the reference's per-node B&B rules
(/root/reference/src/sypha_solver_heuristics.cpp:10-292, /root/reference/src/sypha_solver_bnb.cpp:350-382).

TEST INFRASTRUCTURE - see oracle/__init__.py.  `reference_lib()` loads the reference's OWN code behind the same
signatures (oracle/_ref/libref_heur.so, built by oracle/Makefile where /root/reference exists) for pinning.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None

_I = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is None:
        so = HERE / "_build" / "libheur_oracle.so"
        src = HERE / "heur_oracle.c"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            so.parent.mkdir(exist_ok=True)
            subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", str(src), "-lm", "-o", str(so)], check=True)
        L = C.CDLL(str(so))
        L.oracle_select_branch.restype = C.c_int
        L.oracle_select_branch.argtypes = [C.c_int, _D, _D, C.c_int, C.c_double, C.POINTER(C.c_double)]
        L.oracle_nearest_integer_fixing.restype = C.c_int
        L.oracle_nearest_integer_fixing.argtypes = [C.c_int, C.c_int, _I, _I, _D, _D, _D, _D, C.c_int, _I, _I, C.c_double, _D,
                                                    C.POINTER(C.c_double)]
        L.oracle_dual_guided_cover_repair.restype = C.c_int
        L.oracle_dual_guided_cover_repair.argtypes = [C.c_int, C.c_int, _I, _I, _D, _D, _D, _D, _D, C.c_int, C.c_int, _I, _I,
                                                      C.c_double, _D, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def reference_lib():
    """The reference's own heuristics (oracle/_ref/libref_heur.so) or None where it was not built."""
    so = HERE / "_ref" / "libref_heur.so"
    if not so.exists():
        return None
    L = C.CDLL(str(so))
    L.ref_heuristic.restype = C.c_int
    L.ref_heuristic.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _I, _I, _D, _D, _D, _D, C.c_int, _D, C.c_int, C.c_int,
                                _I, _I, C.c_double, _D, C.POINTER(C.c_double)]
    L.ref_select_branch.restype = C.c_int
    L.ref_select_branch.argtypes = [C.c_int, _D, _D, C.c_int, C.c_double]
    return L


def _dec(decisions):
    var = np.ascontiguousarray([d[0] for d in decisions] or [0], dtype=np.int32)
    fix = np.ascontiguousarray([d[1] for d in decisions] or [0], dtype=np.int32)
    return len(decisions), var, fix


def select_branch(x, obj, n0, tol=1e-6, rule="most_fractional"):
    """-> (column or -1, |x - floor(x + 0.5)| there)"""
    frac = C.c_double()
    j = lib().oracle_select_branch(0 if rule == "most_fractional" else 1, np.ascontiguousarray(x[:n0], dtype=np.float64),
                                   np.ascontiguousarray(obj[:n0], dtype=np.float64), n0, tol, C.byref(frac))
    return j, frac.value


def nearest_integer_fixing(inst, x, decisions=(), tol=1e-6):
    """-> (feasible, objective, solution[n0])   NearestIntegerFixingHeuristic::tryBuild"""
    n0 = inst.n_orig
    sol, obj = np.zeros(n0), C.c_double()
    nd, var, fix = _dec(decisions)
    f = lib().oracle_nearest_integer_fixing(inst.m, n0, inst.offs, inst.inds, inst.vals, inst.c, inst.b,
                                            np.ascontiguousarray(x, dtype=np.float64), nd, var, fix, tol, sol, C.byref(obj))
    return bool(f), obj.value, sol


def dual_guided_cover_repair(inst, x, y, decisions=(), tol=1e-6):
    """-> (feasible, objective, solution[n0], columns added by the repair)   DualGuidedCoverRepairHeuristic::tryBuild"""
    n0 = inst.n_orig
    sol, obj, steps = np.zeros(n0), C.c_double(), C.c_int()
    nd, var, fix = _dec(decisions)
    y = np.ascontiguousarray(y, dtype=np.float64)
    f = lib().oracle_dual_guided_cover_repair(inst.m, n0, inst.offs, inst.inds, inst.vals, inst.c, inst.b,
                                              np.ascontiguousarray(x, dtype=np.float64), y, len(y), nd, var, fix, tol, sol,
                                              C.byref(obj), C.byref(steps))
    return bool(f), obj.value, sol, steps.value


def reference_heuristic(which, inst, x, y, decisions=(), tol=1e-6):
    """The reference's own tryBuild through oracle/_ref/libref_heur.so -> (feasible, objective, solution)."""
    L = reference_lib()
    n0 = inst.n_orig
    sol, obj = np.zeros(n0), C.c_double()
    nd, var, fix = _dec(decisions)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    f = L.ref_heuristic(0 if which == "nearest_integer_fixing" else 1, inst.m, inst.n, n0, inst.offs, inst.inds, inst.vals,
                        inst.c, inst.b, x, len(x), y, len(y), nd, var, fix, tol, sol, C.byref(obj))
    return bool(f), obj.value, sol
