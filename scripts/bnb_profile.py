"""Where does a B&B round spend its host time?  (model build / upload / batched solve / heuristics)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from sypha_b200 import bnb
from sypha_b200.instances import gen_scp
from sypha_b200 import solver as S

slots = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mdl = gen_scp(500, 5000, 0.10, 77)
drv = bnb.BatchedBnb(mdl, slots=slots)
T = {"build": 0.0, "load": 0.0, "solve": 0.0, "heur": 0.0}
orig_build, orig_heur = bnb.build_branch_model, drv.heur.__call__
def tb(*a, **k):
    t = time.perf_counter(); r = orig_build(*a, **k); T["build"] += time.perf_counter() - t; return r
bnb.build_branch_model = tb
orig_copy = S.SyphaNodeSparse.copyModelOnDevice
def tl(self, *a, **k):
    t = time.perf_counter(); r = orig_copy(self, *a, **k); T["load"] += time.perf_counter() - t; return r
S.SyphaNodeSparse.copyModelOnDevice = tl
orig_sb = bnb.solve_batch
def ts(*a, **k):
    l0 = T["load"]; t = time.perf_counter(); r = orig_sb(*a, **k); T["solve"] += time.perf_counter() - t - (T["load"] - l0); return r
bnb.solve_batch = ts
class H:
    def __init__(s, h): s.h = h; s.A = h.A
    def __call__(s, *a, **k):
        t = time.perf_counter(); r = s.h(*a, **k); T["heur"] += time.perf_counter() - t; return r
drv.heur = H(drv.heur)
t0 = time.perf_counter()
st = drv.run(max_nodes=10**9, rounds=12)
tot = time.perf_counter() - t0
print(f"slots {slots}: {st.processed} nodes in {tot*1e3:.0f} ms -> {st.processed/tot:.1f} nodes/s; per node ms:",
      {k: round(1e3 * v / max(st.processed, 1), 2) for k, v in T.items()}, f"lp device {st.lp_device_ms/max(st.processed,1):.2f}")
drv.close()
