#!/usr/bin/env python
"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = None
    out = []
    for r in rows:
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        try:
            samples = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        st = {k: r[i] for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k}
        top_st = sorted(((int(v), k) for k, v in st.items() if v.isdigit() and int(v) > 0), reverse=True)[:3]
        out.append((samples, r[0], r[1].strip(), top_st))
    tot = sum(o[0] for o in out) or 1
    print(f"total samples {tot}")
    for s, line, src, st in sorted(out, reverse=True)[:top]:
        print(f"{s:6d} {100*s/tot:5.1f}%  L{line:>4}: {src[:90]:90s} {' '.join(f'{k[6:]}={v}' for v, k in st)}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
