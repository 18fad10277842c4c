"""CPU: the library's OR-Library reader (sb200_read_scp, host code in the C-ABI library) against the Python reader
and the oracle's restatement of /root/reference/src/model_reader.cpp:90-174."""
import os
from pathlib import Path

import numpy as np
import pytest

from sypha_b200 import instances
from oracle import scp_io


def _write_orlib(path, m, n0, costs, rows, per_line=12):
    """OR-Library layout: numbers wrapped over lines, every line starting with a blank."""
    toks = [str(m), str(n0)] + [str(c) for c in costs]
    for cols in rows:
        toks.append(str(len(cols)))
        toks += [str(int(j) + 1) for j in cols]
    with open(path, "w") as f:
        for i in range(0, len(toks), per_line):
            f.write(" " + " ".join(toks[i:i + per_line]) + " \n")


def _rows_of(mdl):
    return [mdl.inds[mdl.offs[i]:mdl.offs[i + 1] - 1] for i in range(mdl.m)]


def _same(a, b):
    assert (a.m, a.n, a.n_orig, a.nnz) == (b.m, b.n, b.n_orig, b.nnz)
    for k in ("offs", "inds", "vals", "c", "b"):
        x, y = np.asarray(getattr(a, k)), np.asarray(getattr(b, k))
        assert x.dtype == y.dtype and np.array_equal(x, y), k


@pytest.mark.parametrize("shape", [(7, 20, 0.3, 1), (200, 1000, 0.02, 3), (50, 4000, 0.05, 4)])
def test_native_reader_matches_python_and_oracle(tmp_path, shape):
    m, n0, dens, seed = shape
    mdl = instances.gen_scp(m, n0, dens, seed)
    p = tmp_path / "inst.txt"
    _write_orlib(p, m, n0, [int(c) for c in mdl.c[:n0]], _rows_of(mdl))
    nat = instances.read_scp_native(p)
    _same(nat, instances.read_scp(p))
    _same(nat, mdl)
    ora = scp_io.load_scp(p)
    _same(nat, ora)


def test_native_reader_real_costs_and_empty_rows(tmp_path):
    p = tmp_path / "real.txt"
    p.write_text("3 4\n1.5 2 3e1 0.25\n2 1 4\n0\n1 3\n")
    nat = instances.read_scp_native(p)
    assert (nat.m, nat.n, nat.n_orig, nat.nnz) == (3, 7, 4, 6)
    assert nat.c.tolist() == [1.5, 2.0, 30.0, 0.25, 0.0, 0.0, 0.0]
    assert nat.offs.tolist() == [0, 3, 4, 6]
    assert nat.inds.tolist() == [0, 3, 4, 5, 2, 6]
    assert nat.vals.tolist() == [1.0, 1.0, -1.0, -1.0, 1.0, -1.0]
    assert nat.b.tolist() == [1.0, 1.0, 1.0]


@pytest.mark.parametrize("text", ["", "3 4\n1 2 3 4\n2 1 4\n", "2 2\n1 1\n1 3\n1 1\n", "2 2\n1 x\n1 1\n1 2\n", "0 5\n"])
def test_native_reader_rejects_malformed_files(tmp_path, text):
    p = tmp_path / "bad.txt"
    p.write_text(text)
    with pytest.raises(ValueError):
        instances.read_scp_native(p)
    with pytest.raises(ValueError):
        instances.read_scp_native(tmp_path / "missing.txt")


@pytest.mark.skipif(not Path("/root/reference/data/scp41.txt").exists(), reason="reference data not present")
@pytest.mark.parametrize("name", ["scp41", "scpnrh1", "scpclr13"])
def test_native_reader_on_the_reference_instances(name):
    p = f"/root/reference/data/{name}.txt"
    _same(instances.read_scp_native(p), scp_io.load_scp(p))
