"""TEST INFRASTRUCTURE / study (CPU, NumPy): see DESIGN.md.  Not imported by the product, the tests or the bench."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle import scp_io, mehrotra as mo
from sypha_b200 import bnb
from sypha_b200.instances import load_npz

name = sys.argv[1] if len(sys.argv) > 1 else 'scpnre1'
mdl = load_npz(f'/root/repo/tests/golden/{name}.npz')
obj, _ = bnb.greedy_cover(mdl)
red, _ = bnb.reduce_by_incumbent(mdl, obj)
base = scp_io.ScpInstance(red.m, red.n, red.n_orig, red.offs, red.inds, red.vals, red.c, red.b)
P = mo.Params(max_iter=100)

def solve(inst, start=None, stop_after=None):
    return mo.mehrotra(inst.csr(), inst.b, inst.c, inst.n_orig, P, "ne", start=start, stop_after=stop_after)

def child_start(parent_res, scheme, theta, k_back=0, parent_inst=None):
    """start for a child with one more row/column than the parent"""
    if k_back:
        r = solve(parent_inst, stop_after=max(1, parent_res.iterations - k_back))
        x, y, s = r.x, r.y, r.s
    else:
        x, y, s = parent_res.x, parent_res.y, parent_res.s
    x = np.concatenate([x, [0.0]]); s = np.concatenate([s, [0.0]]); y = np.concatenate([y, [0.0]])
    if scheme == 'floor':
        x = np.maximum(x, theta); s = np.maximum(s, theta)
    elif scheme == 'shift':
        x = x + theta; s = s + theta
    elif scheme == 'mu':       # floor products: raise the smaller of each pair so that x_i s_i >= theta^2
        x = np.maximum(x, 1e-12); s = np.maximum(s, 1e-12)
        low = x * s < theta * theta
        xs = np.sqrt(x * s)
        # scale the pair symmetrically up to product theta^2
        f = np.where(low, theta / np.maximum(xs, 1e-300), 1.0)
        x = x * f; s = s * f
    return x, y, s

root = solve(base)
print(f"{name}: root {root.iterations} iterations primal {root.primal:.6f} dual {root.dual:.6f}")
rng = np.random.default_rng(0)
frac = np.abs(root.x[:base.n_orig] - np.round(root.x[:base.n_orig]))
jbr = int(np.argmax(frac))
nodes = [((jbr, 0),), ((jbr, 1),)]
# one more level under the x=1 child
schemes = [('floor', 1e-2, 0), ('floor', 1e-1, 0), ('shift', 1e-2, 0), ('shift', 1e-1, 0), ('mu', 1e-1, 0), ('mu', 3e-1, 0),
           ('floor', 1e-2, 3), ('floor', 1e-1, 3), ('mu', 1e-1, 3), ('floor', 1e-1, 6), ('mu', 3e-1, 6), ('shift', 3e-1, 0), ('shift', 1.0, 0)]
def run_level(parent_inst, parent_res, decs, label):
    for d in decs:
        inst = scp_io.append_branch_rows(parent_inst, [d[-1]])
        cold = solve(inst)
        out = []
        for sch, th, kb in schemes:
            st = child_start(parent_res, sch, th, kb, parent_inst)
            w = solve(inst, start=st)
            out.append(f"{sch}{th:g}/-{kb}:{w.iterations}({(w.primal - cold.primal) / max(1, abs(cold.primal)):+.1e},{w.reason})")
        print(f"{label} dec {d[-1]}: cold {cold.iterations} it primal {cold.primal:.6f} reason {cold.reason} | " + " ".join(out), flush=True)
        yield inst, cold
lvl1 = list(run_level(base, root, nodes, "L1"))
for inst, res in lvl1:
    if res.status != 0: continue
    fr = np.abs(res.x[:base.n_orig] - np.round(res.x[:base.n_orig])); j2 = int(np.argmax(fr))
    list(run_level(inst, res, [((j2, 0),), ((j2, 1),)], "L2"))
