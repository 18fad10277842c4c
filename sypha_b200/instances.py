"""Set-covering LP instances for the host side of the path: the OR-Library text format of the
reference's reader (/root/reference/src/model_reader.cpp:90-174: ``m n``, n costs, then per row
``k idx_1..idx_k`` 1-based; standard form A = [A0 | -I], b = 1, c = [c0; 0]) and the synthetic
generator used by bench.py (SURVEY.md Appendix C shape: k = round(n*density) random columns per row,
no empty columns, integer costs 1..100)."""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class ScpModel:
    m: int
    n: int
    n_orig: int
    offs: np.ndarray     # int32 [m+1]
    inds: np.ndarray     # int32 [nnz]
    vals: np.ndarray     # float64 [nnz]
    c: np.ndarray        # float64 [n]
    b: np.ndarray        # float64 [m]
    name: str = ""

    @property
    def nnz(self):
        return int(self.inds.shape[0])


def _standard_form(m, n, row_cols, row_cnt, costs, name):
    """row_cols: concatenated 0-based column ids, row_cnt[m]: entries per row."""
    cnt = np.asarray(row_cnt, dtype=np.int64)
    offs = np.zeros(m + 1, dtype=np.int64)
    offs[1:] = np.cumsum(cnt + 1)
    inds = np.empty(offs[-1], dtype=np.int32)
    vals = np.ones(offs[-1], dtype=np.float64)
    dst = np.arange(len(row_cols), dtype=np.int64) + np.repeat(np.arange(m, dtype=np.int64), cnt)
    inds[dst] = row_cols
    inds[offs[1:] - 1] = n + np.arange(m, dtype=np.int32)      # surplus column, last in its row
    vals[offs[1:] - 1] = -1.0
    c = np.concatenate([np.asarray(costs, dtype=np.float64), np.zeros(m)])
    return ScpModel(m, n + m, n, offs.astype(np.int32), inds, vals, c, np.ones(m), name)


def read_scp_native(path) -> ScpModel:
    """The library's reader (``sb200_read_scp``, csrc/sb200_io.cu): one pass over the file, arrays written at
    their final size.  Same result as ``read_scp`` below."""
    import ctypes as C
    from . import _lib as L
    lib = L.load()
    mdl = L.sb200_scp_model()
    rc = lib.sb200_read_scp(str(path).encode(), C.byref(mdl))
    if rc != L.SB200_OK:
        raise ValueError(f"sb200_read_scp({path}) failed with code {rc}")
    try:
        out = ScpModel(mdl.m, mdl.n, mdl.n_orig,
                       np.ctypeslib.as_array(mdl.csr_offs, (mdl.m + 1,)).copy(),
                       np.ctypeslib.as_array(mdl.csr_inds, (mdl.nnz,)).copy(),
                       np.ctypeslib.as_array(mdl.csr_vals, (mdl.nnz,)).copy(),
                       np.ctypeslib.as_array(mdl.c, (mdl.n,)).copy(),
                       np.ctypeslib.as_array(mdl.b, (mdl.m,)).copy(), str(path))
    finally:
        lib.sb200_free_scp(C.byref(mdl))
    return out


def read_scp(path) -> ScpModel:
    tok = np.array(open(path).read().split(), dtype=np.int64)
    m, n = int(tok[0]), int(tok[1])
    costs = tok[2:2 + n].astype(np.float64)
    pos, cols, cnt = 2 + n, [], []
    for _ in range(m):
        k = int(tok[pos])
        cols.append(tok[pos + 1:pos + 1 + k] - 1)
        cnt.append(k)
        pos += 1 + k
    return _standard_form(m, n, np.concatenate(cols), cnt, costs, str(path))


def gen_scp(m, n, density, seed) -> ScpModel:
    r = np.random.default_rng(seed)
    k = max(1, int(round(n * density)))
    cols = r.integers(0, n, size=(m, k), dtype=np.int64)
    cols.sort(axis=1)
    keep = np.ones((m, k), dtype=bool)
    keep[:, 1:] = cols[:, 1:] != cols[:, :-1]
    rows = np.repeat(np.arange(m, dtype=np.int64), k).reshape(m, k)[keep]
    cols = cols[keep]
    present = np.zeros(n, dtype=bool)
    present[cols] = True
    empty = np.nonzero(~present)[0]
    if len(empty):
        rows = np.concatenate([rows, r.integers(0, m, len(empty))])
        cols = np.concatenate([cols, empty])
        order = np.lexsort((cols, rows))
        rows, cols = rows[order], cols[order]
    cnt = np.bincount(rows, minlength=m)
    costs = r.integers(1, 101, n).astype(np.float64)
    return _standard_form(m, n, cols, cnt, costs, f"gen_scp({m},{n},{density},{seed})")


def load_npz(path) -> ScpModel:
    """An OR-Library instance from the compact archive the repository ships its instances in (``m``, ``n_orig``,
    ``costs``, ``row_offs``, ``col_inds`` of A0 - tests/golden/*.npz, written by tests/golden/make_golden.py from
    the reference's data/ files); the same standard form as ``read_scp``."""
    z = np.load(path, allow_pickle=False)
    m, n0 = int(z["m"]), int(z["n_orig"])
    offs = z["row_offs"].astype(np.int64)
    return _standard_form(m, n0, z["col_inds"].astype(np.int64), np.diff(offs), z["costs"].astype(np.float64),
                          str(path))


def write_scp(mdl: ScpModel, path) -> None:
    """OR-Library text form of a standard-form model whose last entry per row is the surplus column (the inverse
    of ``read_scp``): what the reference's own CLI reads (model_reader.cpp:90-174)."""
    n0 = mdl.n_orig
    costs = mdl.c[:n0]
    as_int = bool(np.all(costs == np.round(costs)))
    out = [f" {mdl.m} {n0}"]
    for a in range(0, n0, 12):
        out.append(" " + " ".join(str(int(v)) if as_int else repr(float(v)) for v in costs[a:a + 12]))
    for i in range(mdl.m):
        cols = mdl.inds[mdl.offs[i]:mdl.offs[i + 1] - 1] + 1
        out.append(f" {len(cols)}")
        for a in range(0, len(cols), 12):
            out.append(" " + " ".join(map(str, cols[a:a + 12])))
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


def build_standard_form(n_vars: int, row_offs, row_inds, row_vals, row_lb, row_ub, obj, maximize: bool = False,
                        name: str = "model") -> ScpModel:
    """A general row model (lb <= a.x <= ub, x >= 0) -> the standard form the solver takes, through the library's
    ``sb200_build_standard_form`` (csrc/sb200_io.cu; the reference's ``buildStandardForm``, src/sypha_api.cpp:136-250):
    O(nnz), no Variable / Constraint objects.  Returns the arrays ``SyphaNodeSparse.from_csr`` / ``sb200_load_model`` take."""
    import ctypes as C
    from . import _lib as L
    lib = L.load()
    ro = np.ascontiguousarray(row_offs, dtype=np.int32)
    ri = np.ascontiguousarray(row_inds, dtype=np.int32)
    rv = np.ascontiguousarray(row_vals, dtype=np.float64)
    lb = np.ascontiguousarray(row_lb, dtype=np.float64)
    ub = np.ascontiguousarray(row_ub, dtype=np.float64)
    c = np.ascontiguousarray(obj, dtype=np.float64)
    n_rows = len(ro) - 1
    if len(lb) != n_rows or len(ub) != n_rows or len(c) != n_vars or len(ri) != len(rv) or (n_rows >= 0 and len(ri) < ro[-1]):
        raise ValueError("build_standard_form: array lengths do not match the row model")
    PI, PD = C.POINTER(C.c_int), C.POINTER(C.c_double)
    m = L.sb200_row_model(n_vars, n_rows, ro.ctypes.data_as(PI), ri.ctypes.data_as(PI), rv.ctypes.data_as(PD),
                          lb.ctypes.data_as(PD), ub.ctypes.data_as(PD), c.ctypes.data_as(PD), 1 if maximize else 0)
    nr, nc, nz = C.c_int(), C.c_int(), C.c_longlong()
    rc = lib.sb200_standard_form_size(C.byref(m), C.byref(nr), C.byref(nc), C.byref(nz))
    if rc != L.SB200_OK:
        raise ValueError(f"sb200_standard_form_size failed with code {rc}")
    offs = np.empty(nr.value + 1, dtype=np.int32)
    inds = np.empty(nz.value, dtype=np.int32)
    vals = np.empty(nz.value, dtype=np.float64)
    cc = np.empty(nc.value, dtype=np.float64)
    b = np.empty(nr.value, dtype=np.float64)
    rc = lib.sb200_build_standard_form(C.byref(m), offs.ctypes.data_as(PI), inds.ctypes.data_as(PI), vals.ctypes.data_as(PD),
                                       cc.ctypes.data_as(PD), b.ctypes.data_as(PD))
    if rc != L.SB200_OK:
        raise ValueError(f"sb200_build_standard_form failed with code {rc}")
    return ScpModel(nr.value, nc.value, n_vars, offs, inds, vals, cc, b, name)
