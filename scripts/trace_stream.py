#!/usr/bin/env python
"""CUPTI timeline of the continuous batcher (sb200_solve_stream through BatchedBnb.stream_round): are the slots kept busy?"""
import collections, json, sys, tempfile, time
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from sypha_b200 import bnb
from sypha_b200.instances import load_npz
inst = sys.argv[1] if len(sys.argv) > 1 else "scpnre1"
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 128
mdl = load_npz(REPO / "tests" / "golden" / f"{inst}.npz")
drv = bnb.BatchedBnb.with_reference_presolve(mdl, slots=slots)
while len(drv.frontier) < 4 * slots and drv.frontier:
    drv.round()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    n0 = drv.stats.processed
    drv.stream_round(2 * slots)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
nodes = drv.stats.processed - n0
path = Path(tempfile.mkdtemp()) / "trace.json"
prof.export_chrome_trace(str(path))
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
t_begin = min(e["ts"] for e in ev); t_end = max(e["ts"] + e["dur"] for e in ev)
by = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    nm = e["name"].split("(")[-2].split("::")[-1] if "anonymous" in e["name"] else e["name"].split("(")[0].split("::")[-1]
    by[nm][0] += 1; by[nm][1] += e["dur"]
pts = sorted([(e["ts"], 1) for e in ev] + [(e["ts"] + e["dur"], -1) for e in ev])
area, depth, last = 0.0, 0, pts[0][0]
for t, d in pts:
    area += depth * (t - last); depth += d; last = t
print(f"{inst} {slots} slots: {nodes} nodes in {1e3 * wall:.1f} ms host wall, kernels span {(t_end - t_begin) / 1e3:.1f} ms, mean kernels in flight {area / (t_end - t_begin):.1f}")
for nm, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"  {nm[:50]:50s} {c:6d} launches, mean {d / c:9.1f} us")
# gaps between the end of one LP kernel and the start of the next on the same stream
lp = collections.defaultdict(list)
for e in ev:
    if "k_ipm_cta" in e["name"]:
        lp[e["args"].get("stream")].append((e["ts"], e["ts"] + e["dur"]))
gaps = []
for st, lst in lp.items():
    lst.sort()
    gaps += [b[0] - a[1] for a, b in zip(lst, lst[1:])]
if gaps:
    gaps.sort()
    print(f"  LP-to-LP gap on a slot's stream: median {gaps[len(gaps) // 2]:.0f} us, p90 {gaps[int(0.9 * len(gaps))]:.0f} us, max {gaps[-1]:.0f} us ({len(gaps)} gaps)")
drv.close()
