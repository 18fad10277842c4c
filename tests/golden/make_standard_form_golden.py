#!/usr/bin/env python
"""Writes tests/golden/standard_form_cases.json: general row models and the standard form the REFERENCE's own
buildStandardForm (src/sypha_api.cpp:136-250) builds for them, obtained from oracle/_ref/sf_dump_ref (the reference's
translation unit compiled in place by oracle/Makefile).  Run in the container that holds /root/reference:
    make -C oracle all && python tests/golden/make_standard_form_golden.py"""
import json
import math
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
BIN = REPO / "oracle" / "_ref" / "sf_dump_ref"


def fmt(v):
    return "inf" if v == math.inf else "-inf" if v == -math.inf else repr(float(v))


def reference_standard_form(n_vars, rows, lbs, ubs, obj, maximize):
    """rows: [[(var, coef), ...], ...]; obj: [(var, coef), ...] in insertion order"""
    lines = [f"{n_vars} {len(rows)} {1 if maximize else 0} {len(obj)}",
             " ".join(f"{j} {fmt(c)}" for j, c in obj)]
    for r, lb, ub in zip(rows, lbs, ubs):
        lines.append(f"{fmt(lb)} {fmt(ub)} {len(r)} " + " ".join(f"{j} {fmt(c)}" for j, c in r))
    out = subprocess.run([str(BIN)], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
    return json.loads(out)


def random_case(seed):
    rng = np.random.default_rng(seed)
    n_vars, n_rows = 6 + seed, 12 + 3 * seed
    rows, lbs, ubs = [], [], []
    for i in range(n_rows):
        k = int(rng.integers(0, min(n_vars, 5) + 1))
        cols = rng.permutation(n_vars)[:k]
        rows.append([(int(j), float(np.round(rng.normal(), 3))) for j in cols])
        lo = float(np.round(rng.normal(), 2))
        kind = i % 5
        lbs.append(lo if kind in (0, 1, 3) else -math.inf)
        ubs.append(lo if kind == 0 else math.inf if kind in (1, 4) else lo + 1.5 if kind == 3 else lo)
    obj = [(int(j), float(np.round(rng.normal(), 3))) for j in rng.permutation(n_vars)[:max(1, n_vars // 2)]]
    return dict(n_vars=n_vars, rows=rows, lbs=lbs, ubs=ubs, obj=obj, maximize=bool(seed % 2))


if __name__ == "__main__":
    if not BIN.exists():
        sys.exit("oracle/_ref/sf_dump_ref is missing: make -C oracle all")
    cases = []
    for seed in range(6):
        c = random_case(seed)
        ref = reference_standard_form(c["n_vars"], c["rows"], c["lbs"], c["ubs"], c["obj"], c["maximize"])
        c["lbs"] = [fmt(v) for v in c["lbs"]]
        c["ubs"] = [fmt(v) for v in c["ubs"]]
        c["reference"] = ref
        cases.append(c)
    (HERE / "standard_form_cases.json").write_text(json.dumps(cases, indent=0))
    print(f"wrote {len(cases)} cases")
