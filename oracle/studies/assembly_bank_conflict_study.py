"""Study (CPU): shared-memory bank conflicts of the one-block assembly gathers on scpnrh1, before and after the greedy
id order of k_decollide_chunks16.  Half-warp model: 16 lanes, 16 double-wide banks, equal addresses
broadcast.  Not imported by the product, the tests or the bench.   python oracle/studies/assembly_bank_conflict_study.py"""
import numpy as np, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from sypha_b200.instances import load_npz
mdl=load_npz(__import__('pathlib').Path(__file__).resolve().parents[2] / 'tests' / 'golden' / 'scpnrh1.npz')
m,n=mdl.m,mdl.n
rows=[set(mdl.inds[mdl.offs[i]:mdl.offs[i+1]].tolist()) for i in range(m)]
pad=n
rng=np.random.default_rng(0)
def chunks_of(i,k):
    ids=sorted(rows[i]&rows[k])
    out=[]
    for a in range(0,len(ids),8):
        c=ids[a:a+8]; c+= [pad]*(8-len(c)); out.append(c)
    return out
def wave(half):  # half: list of up to 16 chunks (each 8 ids) -> wavefronts for the 8 slot-gathers under half-warp model
    w=0
    for q in range(8):
        addr=set(c[q] for c in half)
        banks={}
        for a in addr: banks[a&15]=banks.get(a&15,0)+1
        w+=max(banks.values()) if banks else 0
    return w
def greedy(half):
    mask=[0]*8; out=[]
    for l,c in enumerate(half):
        o=[None]*8; used=0
        for id_ in c:
            if id_==pad: continue
            bit=1<<(id_&15); best=-1
            for qq in range(8):
                q=(qq+l)&7
                if not (used>>q)&1 and not mask[q]&bit: best=q;break
            if best<0:
                for qq in range(8):
                    q=(qq+l)&7
                    if not (used>>q)&1: best=q;break
            used|=1<<best; o[best]=id_; mask[best]|=bit
        o=[pad if x is None else x for x in o]
        out.append(o)
    return out
tot0=tot1=cnt=0
for i in (200,500,800,999):
    for g in range(0,min(i,160),16):
        ent=[chunks_of(i,k) for k in range(g,min(g+16,i))]
        for j in range(max(len(e) for e in ent)):
            half=[e[j] for e in ent if j<len(e)]
            tot0+=wave(half); tot1+=wave(greedy(half)); cnt+=8
print('avg wavefronts per half-warp gather: before',tot0/cnt,'after',tot1/cnt)


def greedy_counts(half):
    """place every id in the lane's free slot where its bank has been used least (counts instead of a taken/free mask)"""
    cntq = [[0] * 16 for _ in range(8)]
    out = []
    for l, c in enumerate(half):
        o = [None] * 8
        used = 0
        for id_ in c:
            if id_ == pad:
                continue
            b = id_ & 15
            best, bc = -1, 1 << 30
            for qq in range(8):
                q = (qq + l) & 7
                if not (used >> q) & 1 and cntq[q][b] < bc:
                    best, bc = q, cntq[q][b]
            used |= 1 << best
            o[best] = id_
            cntq[best][b] += 1
        out.append([pad if x is None else x for x in o])
    return out


tot2 = cnt2 = 0
for i in (200, 500, 800, 999):
    for g in range(0, min(i, 160), 16):
        ent = [chunks_of(i, k) for k in range(g, min(g + 16, i))]
        for j in range(max(len(e) for e in ent)):
            half = [e[j] for e in ent if j < len(e)]
            tot2 += wave(greedy_counts(half))
            cnt2 += 8
print('count-balancing greedy:', tot2 / cnt2)


def local_search(half, passes=3):
    """count-balancing greedy, then pairwise swaps inside a lane's chunk while they lower the wavefront count"""
    out = [list(c) for c in greedy_counts(half)]
    cntq = [[0] * 16 for _ in range(8)]
    for c in out:
        for q, id_ in enumerate(c):
            if id_ != pad:
                cntq[q][id_ & 15] += 1
    for _ in range(passes):
        improved = False
        for c in out:
            for q1 in range(8):
                for q2 in range(q1 + 1, 8):
                    a, b = c[q1], c[q2]
                    if a == b:
                        continue
                    before = max(cntq[q1]) + max(cntq[q2])
                    if a != pad:
                        cntq[q1][a & 15] -= 1
                        cntq[q2][a & 15] += 1
                    if b != pad:
                        cntq[q2][b & 15] -= 1
                        cntq[q1][b & 15] += 1
                    if max(cntq[q1]) + max(cntq[q2]) < before:
                        c[q1], c[q2] = b, a
                        improved = True
                    else:
                        if a != pad:
                            cntq[q1][a & 15] += 1
                            cntq[q2][a & 15] -= 1
                        if b != pad:
                            cntq[q2][b & 15] += 1
                            cntq[q1][b & 15] -= 1
        if not improved:
            break
    return out


tot3 = cnt3 = 0
for i in (200, 500, 800, 999):
    for g in range(0, min(i, 160), 16):
        ent = [chunks_of(i, k) for k in range(g, min(g + 16, i))]
        for j in range(max(len(e) for e in ent)):
            half = [e[j] for e in ent if j < len(e)]
            tot3 += wave(local_search(half))
            cnt3 += 8
print('count-balancing greedy + local search:', tot3 / cnt3)
