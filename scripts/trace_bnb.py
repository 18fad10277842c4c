#!/usr/bin/env python
"""Kernel timeline of a few B&B rounds (K node LPs in flight on one GPU) through CUPTI (torch.profiler sees every
kernel the process launches, the library's included): where the GPU's time goes when 32 latency-oriented LPs
share it.  Prints per-kernel totals, the union of busy time, the mean number of kernels in flight and the
start-to-start gap statistics of the graph-launched chains.
  python scripts/trace_bnb.py [instance] [slots] [rounds]"""
import collections
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from sypha_b200 import bnb  # noqa: E402
from sypha_b200.instances import load_npz  # noqa: E402

inst = sys.argv[1] if len(sys.argv) > 1 else "scpnre1"
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mdl = load_npz(REPO / "tests" / "golden" / f"{inst}.npz")
drv = bnb.BatchedBnb.with_reference_presolve(mdl, slots=slots)
while len(drv.frontier) < slots and drv.frontier:
    drv.round()
for _ in range(3):
    drv.round()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    n0 = drv.stats.processed
    for _ in range(rounds):
        drv.round()
    torch.cuda.synchronize()
nodes = drv.stats.processed - n0
path = Path(tempfile.mkdtemp()) / "trace.json"
prof.export_chrome_trace(str(path))
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
t0 = min(e["ts"] for e in ev)
t1 = max(e["ts"] + e["dur"] for e in ev)
wall = t1 - t0
by = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    nm = e["name"].split("(")[0].replace("sb200::", "").replace("void ", "")
    by[nm][0] += 1
    by[nm][1] += e["dur"]
# union of busy intervals and mean concurrency
pts = sorted([(e["ts"], 1) for e in ev] + [(e["ts"] + e["dur"], -1) for e in ev])
busy = area = 0.0
depth, last = 0, pts[0][0]
hist = collections.Counter()
for t, d in pts:
    if depth > 0:
        busy += t - last
    area += depth * (t - last)
    hist[min(depth, 64)] += t - last
    depth += d
    last = t
print(f"{inst}: {slots} slots, {rounds} rounds, {nodes} nodes, {len(ev)} kernels in {wall / 1e3:.2f} ms "
      f"({len(ev) / wall:.3f} kernels/us, {wall / max(nodes, 1):.1f} us of wall per node)")
print(f"GPU busy (>= 1 kernel running) {100 * busy / wall:.1f} %, mean kernels in flight {area / wall:.2f}, "
      f"sum of kernel durations {sum(v[1] for v in by.values()) / 1e3:.2f} ms")
print("time with k kernels in flight: " + ", ".join(f"{k}:{100 * v / wall:.0f}%" for k, v in sorted(hist.items()) if v / wall > 0.01))
print(f"{'kernel':60s} {'launches':>8s} {'sum ms':>9s} {'mean us':>9s} {'share of sum':>12s}")
tot = sum(v[1] for v in by.values())
for nm, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{nm[:60]:60s} {c:8d} {d / 1e3:9.2f} {d / c:9.1f} {100 * d / tot:11.1f}%")
drv.close()
