"""Set-covering instance I/O and the synthetic generator (oracle side, CPU only).

TEST INFRASTRUCTURE - see oracle/__init__.py.

* ``read_scp_text``      restates /root/reference/src/model_reader.cpp:90-174 and
                         /root/reference/python/model_importer.py:56-117
                         (OR-Library text: ``m n``, n costs, then per row ``k idx_1..idx_k`` 1-based).
* ``to_standard_form``   restates model_reader.cpp:126-169: A = [A0 | -I], b = 1, c = [c0; 0];
                         the surplus entry is appended *after* the row's own entries.
* ``gen_scp``            SURVEY.md Appendix C (the reference ships no generator).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class ScpInstance:
    """Standard-form LP  min c'x, Ax = b, x >= 0  with A in CSR (int32 indices, float64 values)."""

    __slots__ = ("m", "n", "n_orig", "offs", "inds", "vals", "c", "b", "name")

    def __init__(self, m, n, n_orig, offs, inds, vals, c, b, name=""):
        self.m, self.n, self.n_orig = int(m), int(n), int(n_orig)
        self.offs = np.ascontiguousarray(offs, dtype=np.int32)
        self.inds = np.ascontiguousarray(inds, dtype=np.int32)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.c = np.ascontiguousarray(c, dtype=np.float64)
        self.b = np.ascontiguousarray(b, dtype=np.float64)
        self.name = name

    @property
    def nnz(self):
        return int(self.inds.shape[0])

    def csr(self) -> sp.csr_matrix:
        return sp.csr_matrix((self.vals, self.inds, self.offs), shape=(self.m, self.n))

    def dense(self) -> np.ndarray:
        return self.csr().toarray()


def read_scp_text(path) -> tuple[int, int, np.ndarray, list[np.ndarray]]:
    """Parse an OR-Library SCP file -> (m, n_orig, costs[n_orig], rows[m] of 0-based column ids).

    Token-stream parse, exactly like the reference's fscanf loop (model_reader.cpp:104-150).
    """
    with open(path, "r") as fh:
        tok = np.array(fh.read().split(), dtype=np.int64)
    m, n = int(tok[0]), int(tok[1])
    costs = tok[2:2 + n].astype(np.float64)
    pos = 2 + n
    rows = []
    for _ in range(m):
        k = int(tok[pos])
        rows.append((tok[pos + 1:pos + 1 + k] - 1).astype(np.int32))
        pos += 1 + k
    return m, n, costs, rows


def to_standard_form(m, n_orig, costs, rows, name="") -> ScpInstance:
    """A = [A0 | -I] in CSR with the surplus column last in each row (model_reader.cpp:133-150)."""
    offs = np.zeros(m + 1, dtype=np.int64)
    for i, r in enumerate(rows):
        offs[i + 1] = offs[i] + len(r) + 1
    inds = np.empty(offs[-1], dtype=np.int32)
    vals = np.empty(offs[-1], dtype=np.float64)
    for i, r in enumerate(rows):
        a, e = offs[i], offs[i + 1]
        inds[a:e - 1] = r
        vals[a:e - 1] = 1.0
        inds[e - 1] = n_orig + i
        vals[e - 1] = -1.0
    c = np.concatenate([np.asarray(costs, dtype=np.float64), np.zeros(m)])
    b = np.ones(m)
    return ScpInstance(m, n_orig + m, n_orig, offs, inds, vals, c, b, name)


def load_scp(path, name="") -> ScpInstance:
    m, n, costs, rows = read_scp_text(path)
    return to_standard_form(m, n, costs, rows, name or str(path))


def write_scp_text(inst: ScpInstance, path) -> None:
    """Write the OR-Library text form (``m n``, the costs, then per row ``k idx_1..idx_k`` 1-based) of a
    standard-form instance whose last entry per row is the surplus column - the inverse of ``load_scp``; it is how
    the committed fixtures (tests/golden/*.npz) reach the reference's own binaries (oracle/_ref) on the GPU box,
    where /root/reference/data does not exist.  Integer costs are written as integers, like the originals."""
    n0 = inst.n_orig
    costs = inst.c[:n0]
    with open(path, "w") as fh:
        fh.write(f" {inst.m} {n0}\n")
        as_int = np.all(costs == np.round(costs))
        for a in range(0, n0, 12):
            fh.write(" " + " ".join(str(int(v)) if as_int else repr(float(v)) for v in costs[a:a + 12]) + "\n")
        for i in range(inst.m):
            cols = inst.inds[inst.offs[i]:inst.offs[i + 1] - 1] + 1
            fh.write(f" {len(cols)}\n")
            for a in range(0, len(cols), 12):
                fh.write(" " + " ".join(map(str, cols[a:a + 12])) + "\n")


def gen_scp(m, n, density, seed) -> ScpInstance:
    """Synthetic random SCP, SURVEY.md Appendix C (RNG call order is part of the spec).

    k = max(1, round(n*density)) distinct columns per row; every empty column gets one random
    row; costs uniform integers 1..100; standard form [A0 | -I], b = 1.
    """
    r = np.random.default_rng(seed)
    k = max(1, int(round(n * density)))
    cols = np.concatenate([r.choice(n, k, replace=False) for _ in range(m)])
    rws = np.repeat(np.arange(m), k)
    present = np.zeros(n, dtype=bool)
    present[cols] = True
    empty = np.nonzero(~present)[0]
    if len(empty):
        rws = np.concatenate([rws, r.integers(0, m, len(empty))])
        cols = np.concatenate([cols, empty])
    A0 = sp.csr_matrix((np.ones(len(rws)), (rws, cols)), shape=(m, n))
    A0.sum_duplicates()
    A0.sort_indices()
    costs = r.integers(1, 101, n).astype(np.float64)
    rows = [A0.indices[A0.indptr[i]:A0.indptr[i + 1]] for i in range(m)]
    # values are all 1 (k distinct columns per row, the extra row for an empty column is unique)
    return to_standard_form(m, n, costs, rows, f"gen_scp({m},{n},{density},{seed})")


def gen_scp_fast(m, n, density, seed) -> ScpInstance:
    """Large-shape generator for the bench (50k x 1M would take minutes with per-row ``choice``).

    Same distribution family as ``gen_scp`` (k random columns per row, duplicates within a row
    removed, no empty columns, costs 1..100) but drawn with vectorised ``integers`` so the RNG
    stream differs from Appendix C; only used where no Appendix-C check value exists.
    """
    r = np.random.default_rng(seed)
    k = max(1, int(round(n * density)))
    cols = r.integers(0, n, size=(m, k), dtype=np.int64)
    cols.sort(axis=1)
    keep = np.ones((m, k), dtype=bool)
    keep[:, 1:] = cols[:, 1:] != cols[:, :-1]
    rws = np.repeat(np.arange(m, dtype=np.int64), k).reshape(m, k)[keep]
    cols = cols[keep]
    present = np.zeros(n, dtype=bool)
    present[cols] = True
    empty = np.nonzero(~present)[0]
    if len(empty):
        rws = np.concatenate([rws, r.integers(0, m, len(empty))])
        cols = np.concatenate([cols, empty])
    A0 = sp.csr_matrix((np.ones(len(rws)), (rws, cols)), shape=(m, n))
    A0.sum_duplicates()
    A0.sort_indices()
    costs = r.integers(1, 101, n).astype(np.float64)
    cnt = np.diff(A0.indptr).astype(np.int64)
    offs = np.zeros(m + 1, dtype=np.int64)
    offs[1:] = np.cumsum(cnt + 1)
    inds = np.empty(offs[-1], dtype=np.int32)
    vals = np.ones(offs[-1], dtype=np.float64)
    # scatter the A0 entries, then the surplus column at the end of each row
    dst = np.arange(A0.nnz, dtype=np.int64) + np.repeat(np.arange(m, dtype=np.int64), cnt)
    inds[dst] = A0.indices
    inds[offs[1:] - 1] = n + np.arange(m, dtype=np.int32)
    vals[offs[1:] - 1] = -1.0
    c = np.concatenate([costs, np.zeros(m)])
    return ScpInstance(m, n + m, n, offs, inds, vals, c, np.ones(m),
                       f"gen_scp_fast({m},{n},{density},{seed})")


def append_branch_rows(inst: ScpInstance, decisions) -> ScpInstance:
    """Node model = base + one row per branching decision, restating
    /root/reference/src/sypha_solver_bnb.cpp:418-468 (``build_branch_model``):
    row = (fix==0 ? -1 : +1) at ``var``, -1 at a new slack column; rhs = fix; cost 0.
    ``decisions`` is a list of (var_index, fix_value in {0,1})."""
    k = len(decisions)
    if k == 0:
        return inst
    offs = np.concatenate([inst.offs.astype(np.int64),
                           inst.offs[-1] + 2 * np.arange(1, k + 1, dtype=np.int64)])
    inds = np.empty(inst.nnz + 2 * k, dtype=np.int32)
    vals = np.empty(inst.nnz + 2 * k, dtype=np.float64)
    inds[:inst.nnz] = inst.inds
    vals[:inst.nnz] = inst.vals
    rhs = []
    for r, (var, fix) in enumerate(decisions):
        p = inst.nnz + 2 * r
        inds[p], vals[p] = var, (-1.0 if fix == 0 else 1.0)
        inds[p + 1], vals[p + 1] = inst.n + r, -1.0
        rhs.append(float(fix))
    c = np.concatenate([inst.c, np.zeros(k)])
    b = np.concatenate([inst.b, np.array(rhs)])
    return ScpInstance(inst.m + k, inst.n + k, inst.n_orig, offs, inds, vals, c, b,
                       inst.name + f"+{k}br")
