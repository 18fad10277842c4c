"""CPU: the library's general-row-model -> standard-form builder (sb200_build_standard_form, host code in
csrc/sb200_io.cu) against the oracle's restatement of the reference's buildStandardForm (src/sypha_api.cpp:136-250):
every row kind (equality, >=, <=, range, free), maximisation, insertion order of the coefficients, and the
set-covering case, where it must give the model the SCP reader gives."""
import math

import numpy as np
import pytest

from oracle import standard_form as osf
from sypha_b200 import _lib as L
from sypha_b200.instances import build_standard_form, gen_scp

INF = math.inf


def _random_model(rng, n_vars, n_rows):
    rows, lbs, ubs = [], [], []
    for i in range(n_rows):
        k = int(rng.integers(0, min(n_vars, 6) + 1))
        cols = rng.permutation(n_vars)[:k]                     # insertion order, not sorted
        rows.append([(int(j), float(np.round(rng.normal(), 3))) for j in cols])
        kind = i % 5
        lo = float(np.round(rng.normal(), 2))
        if kind == 0:
            lbs.append(lo); ubs.append(lo)                     # equality
        elif kind == 1:
            lbs.append(lo); ubs.append(INF)                    # >=
        elif kind == 2:
            lbs.append(-INF); ubs.append(lo)                   # <=
        elif kind == 3:
            lbs.append(lo); ubs.append(lo + 1.5)               # range -> two rows
        else:
            lbs.append(-INF); ubs.append(INF)                  # free
    obj = {int(j): float(np.round(rng.normal(), 3)) for j in rng.permutation(n_vars)[:max(1, n_vars // 2)]}
    return rows, lbs, ubs, obj


def _as_arrays(n_vars, rows, obj):
    offs = np.zeros(len(rows) + 1, dtype=np.int32)
    for i, r in enumerate(rows):
        offs[i + 1] = offs[i] + len(r)
    inds = np.array([j for r in rows for j, _ in r], dtype=np.int32)
    vals = np.array([v for r in rows for _, v in r], dtype=np.float64)
    c = np.zeros(n_vars)
    for j, v in obj.items():
        c[j] = v
    return offs, inds, vals, c


@pytest.mark.parametrize("seed,maximize", [(0, False), (1, True), (2, False), (3, True)])
def test_builder_matches_the_restatement_of_the_reference(seed, maximize):
    rng = np.random.default_rng(seed)
    n_vars, n_rows = 9 + seed, 23 + 5 * seed
    rows, lbs, ubs, obj = _random_model(rng, n_vars, n_rows)
    want = osf.build_standard_form(n_vars, rows, lbs, ubs, obj, maximize)
    offs, inds, vals, c = _as_arrays(n_vars, rows, obj)
    got = build_standard_form(n_vars, offs, inds, vals, lbs, ubs, c, maximize)
    nrows, ncols, w_offs, w_inds, w_vals, w_obj, w_rhs = want
    assert (got.m, got.n, got.n_orig) == (nrows, ncols, n_vars)
    assert got.offs.tolist() == w_offs and got.inds.tolist() == w_inds
    assert np.array_equal(got.vals, np.array(w_vals)) and np.array_equal(got.c, np.array(w_obj))
    assert np.array_equal(got.b, np.array(w_rhs))
    # one surplus column per inequality row, each used once, with -1
    surplus = got.inds[got.inds >= n_vars]
    assert sorted(surplus.tolist()) == list(range(n_vars, ncols))
    assert np.all(got.vals[got.inds >= n_vars] == -1.0)


def test_set_covering_rows_give_the_readers_model():
    """a.x >= 1 per row: [A0 | -I], b = 1, c = [c0; 0] - the model sb200_read_scp / the generator produce."""
    mdl = gen_scp(12, 40, 0.2, 5)
    n0 = mdl.n_orig
    rows = []
    for i in range(mdl.m):
        cols = mdl.inds[mdl.offs[i]:mdl.offs[i + 1]]
        rows.append([(int(j), 1.0) for j in cols if j < n0])
    offs, inds, vals, _ = _as_arrays(n0, rows, {})
    got = build_standard_form(n0, offs, inds, vals, [1.0] * mdl.m, [INF] * mdl.m, mdl.c[:n0])
    assert (got.m, got.n, got.n_orig) == (mdl.m, mdl.n, n0)
    assert np.array_equal(got.offs, mdl.offs) and np.array_equal(got.inds, mdl.inds)
    assert np.array_equal(got.vals, mdl.vals) and np.array_equal(got.c, mdl.c) and np.array_equal(got.b, mdl.b)


def test_empty_and_malformed_models():
    import ctypes as C
    got = build_standard_form(3, [0], [], [], [], [], [1.0, 2.0, 3.0])
    assert (got.m, got.n, got.nnz) == (0, 3, 0) and got.c.tolist() == [1.0, 2.0, 3.0]
    got = build_standard_form(2, [0, 0, 1], [1], [4.0], [-INF, 2.0], [INF, 2.0], [0.0, 0.0])     # an empty free row, an equality
    assert got.offs.tolist() == [0, 0, 1] and got.b.tolist() == [0.0, 2.0] and got.n == 2
    with pytest.raises(ValueError):
        build_standard_form(2, [0, 1], [5], [1.0], [0.0], [INF], [0.0, 0.0])                  # variable out of range
    lib = L.load()
    nr, nc, nz = C.c_int(), C.c_int(), C.c_longlong()
    assert lib.sb200_standard_form_size(None, C.byref(nr), C.byref(nc), C.byref(nz)) == L.SB200_ERR_INVALID
    bad = L.sb200_row_model(2, -1, None, None, None, None, None, None, 0)
    assert lib.sb200_standard_form_size(C.byref(bad), C.byref(nr), C.byref(nc), C.byref(nz)) == L.SB200_ERR_INVALID


# ---- pinned to the reference's OWN builder ---------------------------------------------------------------------------
def _bound(t):
    return INF if t == "inf" else -INF if t == "-inf" else float(t)


def _check_against_reference(case, ref):
    n_vars = case["n_vars"]
    rows = [[(int(j), float(c)) for j, c in r] for r in case["rows"]]
    lbs, ubs = [_bound(t) if isinstance(t, str) else t for t in case["lbs"]], [_bound(t) if isinstance(t, str) else t for t in case["ubs"]]
    obj = {}
    for j, c in case["obj"]:
        obj[int(j)] = float(c)
    want = osf.build_standard_form(n_vars, rows, lbs, ubs, obj, case["maximize"])
    nrows, ncols, w_offs, w_inds, w_vals, w_obj, w_rhs = want
    # the restatement against the reference's code
    assert (nrows, ncols, n_vars) == (ref["nrows"], ref["ncols"], ref["ncols_original"])
    assert w_offs == [int(v) for v in ref["offs"]] and w_inds == [int(v) for v in ref["inds"]]
    assert w_vals == ref["vals"] and w_obj == ref["obj"] and w_rhs == ref["rhs"]
    # the library's builder against the reference's code
    offs, inds, vals, c = _as_arrays(n_vars, rows, obj)
    got = build_standard_form(n_vars, offs, inds, vals, lbs, ubs, c, case["maximize"])
    assert (got.m, got.n) == (ref["nrows"], ref["ncols"])
    assert got.offs.tolist() == [int(v) for v in ref["offs"]] and got.inds.tolist() == [int(v) for v in ref["inds"]]
    assert got.vals.tolist() == ref["vals"] and got.c.tolist() == ref["obj"] and got.b.tolist() == ref["rhs"]


def test_fixtures_written_by_the_references_own_builder():
    """tests/golden/standard_form_cases.json: outputs of SolverImpl::buildStandardForm itself (oracle/_ref/sf_dump_ref =
    the reference's sypha_api.cpp compiled in place), written by tests/golden/make_standard_form_golden.py."""
    import json
    from conftest import GOLDEN
    cases = json.load(open(GOLDEN / "standard_form_cases.json"))
    assert len(cases) >= 6
    for case in cases:
        _check_against_reference(case, case["reference"])


def test_live_against_the_reference_build_when_it_is_here():
    """Fresh random models through oracle/_ref/sf_dump_ref (present where `make -C oracle all` has run)."""
    import importlib.util
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_sf", GOLDEN / "make_standard_form_golden.py")
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    if not mk.BIN.exists():
        pytest.skip("oracle/_ref/sf_dump_ref not built (make -C oracle all)")
    for seed in (11, 12, 13, 14):
        case = mk.random_case(seed)
        try:
            ref = mk.reference_standard_form(case["n_vars"], case["rows"], case["lbs"], case["ubs"], case["obj"], case["maximize"])
        except (OSError, __import__('subprocess').CalledProcessError) as e:   # e.g. a box where the binary does not start
            pytest.skip(f"sf_dump_ref does not start here: {e}")
        _check_against_reference(case, ref)
