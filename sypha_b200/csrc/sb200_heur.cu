// sb200_heur.cu - the per-node combinatorial step of the batched B&B, on the device.
//
// After a node's LP relaxation the reference's driver runs, on the host and one node at a time,
//   * the branching-variable rule MostFractionalSelector
//     (/root/reference/src/sypha_solver_heuristics.cpp:10-30), and
//   * its integer heuristics NearestIntegerFixingHeuristic (:53-110) and the cover repair of
//     DualGuidedCoverRepairHeuristic (:112-292)
// on a host copy of the LP point (/root/reference/src/sypha_solver_bnb_driver.cpp:861-1005).  With the LP
// itself at a few milliseconds, that host step (about 2 ms per node in NumPy) bounds nodes/s once 16+ node
// LPs run concurrently.  Here the plain versions of those rules run as ONE single-CTA
// kernel per node on the node's own stream, right behind the LP, reading the LP point where the solver left
// it; K nodes of a window overlap on the device and 48 bytes per node return to the host.
//
// What the kernel computes (identical, tie-breaks included, to sypha_b200/bnb.py::CoverHeuristic, which the
// tests keep as the checker):
//   1. x_j = [x_lp_j >= 0.5] for the original columns, columns fixed to 0 by the node's decisions excluded;
//   2. greedy repair: while a row is uncovered take the usable column with the smallest cost per newly
//      covered row (first index on ties); gains are kept up to date incrementally;
//   3. redundant columns are dropped, dearest first (ties: lowest index first);
//   4. branching variable = most fractional original column (first index on ties), |x - rint(x)|;
//   5. c . rint(x), the incumbent offer when the LP point is integral.
// All integer work is exact; the two objectives are sums of products cost x {0,1} in a fixed tree order.
#include "sb200_kernels.cuh"
#include "sb200_heur.cuh"

namespace sb200 {

namespace {

constexpr int HT = 1024;          // threads of the single CTA
constexpr int HW = HT / 32;

struct MinKey
{
    double v;
    int j;
};

__device__ __forceinline__ bool key_less(double va, int ja, double vb, int jb)
{
    return va < vb || (va == vb && ja < jb);
}

// block-wide lexicographic minimum of (v, j); result valid in every thread
__device__ MinKey block_min(double v, int j, MinKey *sred)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ov = __shfl_xor_sync(full, v, o);
        const int oj = __shfl_xor_sync(full, j, o);
        if (key_less(ov, oj, v, j)) { v = ov; j = oj; }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                       // sred may still be read from the previous call
    if (l == 0) { sred[w].v = v; sred[w].j = j; }
    __syncthreads();
    v = sred[l].v;                         // HW == 32: one entry per lane
    j = sred[l].j;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ov = __shfl_xor_sync(full, v, o);
        const int oj = __shfl_xor_sync(full, j, o);
        if (key_less(ov, oj, v, j)) { v = ov; j = oj; }
    }
    return MinKey{v, j};
}

__device__ double block_sum(double v, double *sred)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sred[w] = v;
    __syncthreads();
    v = sred[l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
    return v;
}

// state bits of a column
constexpr unsigned char ST_X = 1, ST_BANNED = 2;

__global__ void __launch_bounds__(HT, 1) k_node_heuristics(HeurArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *cover = reinterpret_cast<int *>(smem_raw);                 // [m0]
    int *gain = cover + a.m0;                                       // [n0]
    unsigned char *state = reinterpret_cast<unsigned char *>(gain + a.n0);   // [n0]
    __shared__ MinKey sred[HW];
    __shared__ double sredd[HW];
    __shared__ int s_unc, s_chosen, s_steps;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = a.m0, n0 = a.n0;

    // ---- 4./5. branching variable and the rounded objective (on the LP point itself) -------------------
    {
        double best = -1.0;
        int bj = n0;
        double racc = 0.0;
        for (int j = tid; j < n0; j += HT)
        {
            const double x = a.x_lp[j];
            const double r = rint(x);                               // half to even, as numpy.round
            const double f = fabs(x - r);
            if (f > best) { best = f; bj = j; }                     // j ascends: first index kept on ties
            racc += a.c[j] * r;
        }
        const MinKey k = block_min(-best, bj, sred);
        const double rsum = block_sum(racc, sredd);
        if (tid == 0)
        {
            a.out->branch_var = k.j < n0 ? k.j : -1;
            a.out->branch_frac = -k.v;
            a.out->rounded_obj = rsum;
        }
    }

    // ---- 1. rounding ---------------------------------------------------------------------------------
    for (int j = tid; j < n0; j += HT) state[j] = a.x_lp[j] >= 0.5 ? ST_X : 0;
    for (int i = tid; i < m0; i += HT) cover[i] = 0;
    if (tid == 0) { s_unc = 0; s_chosen = 0; s_steps = 0; }
    __syncthreads();
    for (int r = tid; r < a.k; r += HT)
        if (a.coef[r] < 0.0)                                        // decision "var = 0" (bnb.cpp:453-468)
        {
            const int v = a.var[r];
            if (v >= 0 && v < n0) state[v] = ST_BANNED;
        }
    __syncthreads();
    // cover_i = number of chosen columns in row i.  Through the CSC lists of the CHOSEN columns (a few dozen), not
    // through all rows of the CSR copy: one CTA reading the whole matrix twice (cover counts by rows, gains by
    // columns) was most of this kernel's 350 us.
    for (int j = tid; j < n0; j += HT)
        if (state[j] & ST_X) a.list[atomicAdd(&s_chosen, 1)] = j;
    __syncthreads();
    {
        const int nsel = s_chosen;
        for (int q = warp; q < nsel; q += HW)
        {
            const int j = a.list[q];
            for (int p = a.col_ptr[j] + lane; p < a.col_ptr[j + 1]; p += 32)
            {
                const int i = a.col_rows[p];
                if (i < m0) atomicAdd(&cover[i], 1);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < m0; i += HT)
        if (cover[i] == 0) atomicAdd(&s_unc, 1);
    if (tid == 0) s_chosen = 0;                 // reused by the redundancy pass
    __syncthreads();

    // ---- 2. greedy repair ------------------------------------------------------------------------------
    int feasible = 1;
    if (s_unc > 0)
    {
        // gain_j = uncovered rows column j would cover: every uncovered row (few) adds one to the columns it holds
        for (int j = tid; j < n0; j += HT) gain[j] = 0;
        __syncthreads();
        for (int i = warp; i < m0; i += HW)
            if (cover[i] == 0)
                for (int p = a.row_ptr[i] + lane; p < a.row_ptr[i + 1]; p += 32)
                {
                    const int j = a.row_cols[p];
                    if (j < n0) atomicAdd(&gain[j], 1);
                }
        __syncthreads();
        while (true)
        {
            if (s_unc == 0) break;                                  // uniform: read after a barrier
            double best = DBL_MAX;
            int bj = n0;
            for (int j = tid; j < n0; j += HT)
                if (state[j] == 0 && gain[j] > 0)
                {
                    const double sc = a.c[j] / (double)gain[j];
                    if (sc < best) { best = sc; bj = j; }
                }
            const MinKey k = block_min(best, bj, sred);
            if (k.j >= n0) { feasible = 0; break; }                 // no usable column covers anything new
            const int jc = k.j;
            if (tid == 0) { state[jc] = ST_X; ++s_steps; }
            // rows of the chosen column: warp per row; a newly covered row lowers the gain of its columns
            const int c0 = a.col_ptr[jc], c1 = a.col_ptr[jc + 1];
            for (int q = c0 + warp; q < c1; q += HW)
            {
                const int i = a.col_rows[q];
                if (i >= m0) continue;
                int old = 0;
                if (lane == 0) { old = cover[i]; cover[i] = old + 1; }   // a row appears once per column
                old = __shfl_sync(0xffffffffu, old, 0);
                if (old == 0)
                {
                    if (lane == 0) atomicSub(&s_unc, 1);
                    for (int p = a.row_ptr[i] + lane; p < a.row_ptr[i + 1]; p += 32)
                    {
                        const int j = a.row_cols[p];
                        if (j < n0) atomicSub(&gain[j], 1);
                    }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- 3. drop redundant columns, dearest first ------------------------------------------------------
    if (feasible)
    {
        for (int j = tid; j < n0; j += HT)
            if (state[j] & ST_X) a.list[atomicAdd(&s_chosen, 1)] = j;
        __syncthreads();
        const int nc = s_chosen;
        for (int p = tid; p < nc; p += HT)
        {   // rank by counting under the order (cost descending, index ascending)
            const int jp = a.list[p];
            const double cp = a.c[jp];
            int rank = 0;
            for (int q = 0; q < nc; ++q)
            {
                const int jq = a.list[q];
                const double cq = a.c[jq];
                rank += (cq > cp || (cq == cp && jq < jp)) ? 1 : 0;
            }
            a.sorted[rank] = jp;
        }
        __syncthreads();
        if (warp == 0)
        {
            for (int t = 0; t < nc; ++t)
            {
                const int j = a.sorted[t];
                const int c0 = a.col_ptr[j], c1 = a.col_ptr[j + 1];
                bool red = true;
                for (int q = c0 + lane; q < c1; q += 32)
                {
                    const int i = a.col_rows[q];
                    if (i < m0 && cover[i] < 2) red = false;
                }
                if (__all_sync(0xffffffffu, red))
                {
                    for (int q = c0 + lane; q < c1; q += 32)
                    {
                        const int i = a.col_rows[q];
                        if (i < m0) cover[i] -= 1;
                    }
                    if (lane == 0) state[j] = 0;
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }

    // ---- objective and the cover itself ------------------------------------------------------------------
    double acc = 0.0;
    int cnt = 0;
    for (int j = tid; j < n0; j += HT)
    {
        const unsigned char sel = (state[j] & ST_X) ? 1 : 0;
        a.cover_x[j] = sel;
        if (sel) { acc += a.c[j]; ++cnt; }
    }
    const double obj = block_sum(acc, sredd);
    const double nsel = block_sum((double)cnt, sredd);
    if (tid == 0)
    {
        a.out->feasible = feasible;
        a.out->cover_obj = feasible ? obj : DBL_MAX;
        a.out->n_chosen = (int)nsel;
        a.out->repair_steps = s_steps;
        a.out->nif_feasible = 0;                 // the plain rules have no separate no-repair rounding
        a.out->nif_obj = DBL_MAX;
        if (a.host_flag)
        {
            __threadfence_system();
            *a.host_flag = a.flag_value;
        }
    }
}


// =====================================================================================================
// The reference's rules, exactly (rules = SB200_HEUR_REFERENCE): collect_fractional_candidates + the configured
// selector (sypha_solver_bnb.cpp:368-382, sypha_solver_heuristics.cpp:10-51), NearestIntegerFixingHeuristic
// (:53-110: rounding, decisions imposed, feasible or nothing - no repair) and DualGuidedCoverRepairHeuristic
// (:112-292: start from the fixings and the x_j >= 1 - tol columns; while a row is uncovered add the column with the
// largest (uncovered rows + sum of max(0, y_i) over them) / max(1e-9, c_j), first index on ties; then drop
// redundant non-fixed columns dearest first).  Checked for identity (tests/test_gpu_bnb.py) against the CPU restatement of these rules, which is pinned
// to the reference's own translation unit.  For set-covering models: unit coefficients in the base rows, rhs 1
// (anything else is reported through repair_steps = -1 and no cover).  The dual sums are accumulated over a
// column's uncovered rows in ascending row order, the order of the reference's scan, so the scores - and with
// them every choice - are bit-identical.
constexpr unsigned char ST_FIXED1 = 4, ST_NIF = 8;

__device__ __forceinline__ void node_heuristics_ref_body(const HeurArgs &a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ypos = reinterpret_cast<double *>(smem_raw);                 // [m0] max(0, y_i)
    int *cover = reinterpret_cast<int *>(ypos + a.m0);                   // [m0]
    int *ug = cover + a.m0;                                              // [n0] uncovered rows a column would cover
    unsigned char *state = reinterpret_cast<unsigned char *>(ug + a.n0); // [n0]
    unsigned char *dirty = state + a.n0;                                 // [n0] score must be recomputed
    __shared__ MinKey sred[HW];
    __shared__ double sredd[HW];
    __shared__ int s_unc, s_chosen, s_steps, s_bad;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = a.m0, n0 = a.n0;
    const double tol = a.tol;

    if (tid == 0) { s_unc = 0; s_chosen = 0; s_steps = 0; s_bad = 0; }
    for (int i = tid; i < m0; i += HT)
    {
        const double y = a.y_lp[i];
        ypos[i] = y > 0.0 ? y : 0.0;
        cover[i] = 0;
        if (a.b[i] != 1.0) s_bad = 1;
    }

    // ---- branching variable, rounded objective, the rounding of NearestIntegerFixing ---------------------
    {
        double best = -DBL_MAX;
        int bj = n0;
        double racc = 0.0;
        for (int j = tid; j < n0; j += HT)
        {
            const double v = a.x_lp[j];
            const double nearest = floor(v + 0.5);
            const double f = fabs(v - nearest);
            if (f > tol || nearest < -tol || nearest > 1.0 + tol)
            {
                const double score = a.branch_rule == 0 ? f : a.c[j];
                if (score > best) { best = score; bj = j; }         // j ascends: first index kept on ties
            }
            const double r01 = nearest < 0.0 ? 0.0 : (nearest > 1.0 ? 1.0 : nearest);
            racc += a.c[j] * r01;
            state[j] = r01 > 0.5 ? ST_NIF : 0;
        }
        const MinKey k = block_min(-best, bj, sred);
        const double rsum = block_sum(racc, sredd);
        if (tid == 0)
        {
            const int bv = k.j < n0 ? k.j : -1;
            a.out->branch_var = bv;
            a.out->branch_frac = bv >= 0 ? fabs(a.x_lp[bv] - floor(a.x_lp[bv] + 0.5)) : 0.0;
            a.out->rounded_obj = rsum;
        }
    }
    __syncthreads();
    for (int r = tid; r < a.k; r += HT)
    {
        const int v = a.var[r];
        if (v >= 0 && v < n0) state[v] = a.coef[r] < 0.0 ? 0 : ST_NIF;     // decisions imposed (:77-83)
    }
    __syncthreads();

    // ---- NearestIntegerFixing: feasible as it stands, or nothing ----------------------------------------
    for (int j = tid; j < n0; j += HT)
        if (state[j] & ST_NIF) a.list[atomicAdd(&s_chosen, 1)] = j;
    __syncthreads();
    {
        const int nsel = s_chosen;
        for (int q = warp; q < nsel; q += HW)
        {
            const int j = a.list[q];
            for (int p = a.col_ptr[j] + lane; p < a.col_ptr[j + 1]; p += 32)
            {
                const int i = a.col_rows[p];
                if (i < m0)
                {
                    atomicAdd(&cover[i], 1);
                    if (a.col_vals[p] != 1.0) s_bad = 1;
                }
            }
        }
    }
    __syncthreads();
    {
        int unc = 0;
        double acc = 0.0;
        for (int i = tid; i < m0; i += HT) unc += cover[i] == 0 ? 1 : 0;
        for (int j = tid; j < n0; j += HT)
        {
            const unsigned char sel = (state[j] & ST_NIF) ? 1 : 0;
            a.nif_x[j] = sel;
            if (sel) acc += a.c[j];
        }
        const double tot_unc = block_sum((double)unc, sredd);
        const double obj = block_sum(acc, sredd);
        if (tid == 0)
        {
            a.out->nif_feasible = tot_unc == 0.0 ? 1 : 0;
            a.out->nif_obj = tot_unc == 0.0 ? obj : DBL_MAX;
            s_chosen = 0;
        }
    }
    __syncthreads();

    // ---- DualGuidedCoverRepair: start ------------------------------------------------------------------
    for (int i = tid; i < m0; i += HT) cover[i] = 0;
    for (int j = tid; j < n0; j += HT)
    {
        state[j] = a.x_lp[j] >= 1.0 - tol ? ST_X : 0;
        ug[j] = 0;
        dirty[j] = 1;
    }
    __syncthreads();
    for (int r = tid; r < a.k; r += HT)
    {
        const int v = a.var[r];
        if (v >= 0 && v < n0) state[v] = a.coef[r] < 0.0 ? ST_BANNED : (ST_X | ST_FIXED1);
    }
    __syncthreads();
    for (int j = tid; j < n0; j += HT)
        if (state[j] & ST_X) a.list[atomicAdd(&s_chosen, 1)] = j;
    __syncthreads();
    {
        const int nsel = s_chosen;
        for (int q = warp; q < nsel; q += HW)
        {
            const int j = a.list[q];
            for (int p = a.col_ptr[j] + lane; p < a.col_ptr[j + 1]; p += 32)
            {
                const int i = a.col_rows[p];
                if (i < m0)
                {
                    atomicAdd(&cover[i], 1);
                    if (a.col_vals[p] != 1.0) s_bad = 1;
                }
            }
        }
    }
    __syncthreads();
    for (int i = warp; i < m0; i += HW)
        if (cover[i] == 0)
        {
            if (lane == 0) atomicAdd(&s_unc, 1);
            for (int p = a.row_ptr[i] + lane; p < a.row_ptr[i + 1]; p += 32)
            {
                const int j = a.row_cols[p];
                if (j < n0)
                {
                    atomicAdd(&ug[j], 1);
                    if (a.row_vals[p] != 1.0) s_bad = 1;
                }
            }
        }
    if (tid == 0) s_chosen = 0;
    __syncthreads();

    // ---- repair -----------------------------------------------------------------------------------------
    // A column's score only changes when one of ITS rows gets covered: scores are cached (a.score) and recomputed
    // for the columns of the rows the last pick covered ("dirty"), from scratch and in row order, so every score is
    // the number the reference's full rescan would produce.  (The first version rescanned every column every round:
    // 8 ms per node at scpnrg size, more than the node's LP.)
    int feasible = 1;
    while (true)
    {
        if (s_unc == 0) break;                                      // uniform: read after a barrier
        double best = -DBL_MAX;
        int bj = n0;
        for (int j = tid; j < n0; j += HT)
            if (state[j] == 0 && ug[j] > 0)
            {
                double score;
                if (dirty[j])
                {
                    double dg = 0.0;
                    const int pe = a.col_ptr[j + 1];
                    int p = a.col_ptr[j];
                    for (; p + 4 <= pe; p += 4)
                    {   // four row ids requested together (the walk is a chain of dependent loads otherwise: ncu had
                        // this kernel at 12 long-scoreboard stall cycles per issue); the ADDS stay in row order
                        const int i0 = __ldg(a.col_rows + p), i1 = __ldg(a.col_rows + p + 1), i2 = __ldg(a.col_rows + p + 2),
                                  i3 = __ldg(a.col_rows + p + 3);
                        const bool u0 = i0 < m0 && cover[i0] == 0, u1 = i1 < m0 && cover[i1] == 0,
                                   u2 = i2 < m0 && cover[i2] == 0, u3 = i3 < m0 && cover[i3] == 0;
                        const double y0 = u0 ? ypos[i0] : 0.0, y1 = u1 ? ypos[i1] : 0.0, y2 = u2 ? ypos[i2] : 0.0,
                                     y3 = u3 ? ypos[i3] : 0.0;
                        if (u0) dg = __dadd_rn(dg, y0);
                        if (u1) dg = __dadd_rn(dg, y1);
                        if (u2) dg = __dadd_rn(dg, y2);
                        if (u3) dg = __dadd_rn(dg, y3);
                    }
                    for (; p < pe; ++p)                                       // rows ascending: the reference's order
                    {
                        const int i = __ldg(a.col_rows + p);
                        if (i < m0 && cover[i] == 0) dg = __dadd_rn(dg, ypos[i]);
                    }
                    const double cost = a.c[j] > 1e-9 ? a.c[j] : 1e-9;
                    score = __ddiv_rn(__dadd_rn((double)ug[j], dg), cost);
                    a.score[j] = score;
                    dirty[j] = 0;
                }
                else
                    score = a.score[j];
                if (score > best) { best = score; bj = j; }
            }
        const MinKey k = block_min(-best, bj, sred);
        if (k.j >= n0) { feasible = 0; break; }                     // nothing usable covers an uncovered row (:215-243 finds nothing either)
        const int jc = k.j;
        if (tid == 0) { state[jc] = ST_X; ++s_steps; }
        const int c0 = a.col_ptr[jc], c1 = a.col_ptr[jc + 1];
        for (int q = c0 + warp; q < c1; q += HW)
        {
            const int i = a.col_rows[q];
            if (i >= m0) continue;
            int old = 0;
            if (lane == 0) { old = cover[i]; cover[i] = old + 1; }
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == 0)
            {
                if (lane == 0) atomicSub(&s_unc, 1);
                for (int p = a.row_ptr[i] + lane; p < a.row_ptr[i + 1]; p += 32)
                {
                    const int j = a.row_cols[p];
                    if (j < n0)
                    {
                        atomicSub(&ug[j], 1);
                        dirty[j] = 1;
                    }
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();

    // ---- redundant non-fixed columns, dearest first (equal costs: lowest index first) -------------------------
    if (feasible)
    {
        for (int j = tid; j < n0; j += HT)
            if ((state[j] & ST_X) && !(state[j] & ST_FIXED1)) a.list[atomicAdd(&s_chosen, 1)] = j;
        __syncthreads();
        const int nc = s_chosen;
        for (int p = tid; p < nc; p += HT)
        {
            const int jp = a.list[p];
            const double cp = a.c[jp];
            int rank = 0;
            for (int q = 0; q < nc; ++q)
            {
                const int jq = a.list[q];
                const double cq = a.c[jq];
                rank += (cq > cp || (cq == cp && jq < jp)) ? 1 : 0;
            }
            a.sorted[rank] = jp;
        }
        __syncthreads();
        if (warp == 0)
        {
            for (int t = 0; t < nc; ++t)
            {
                const int j = a.sorted[t];
                const int c0 = a.col_ptr[j], c1 = a.col_ptr[j + 1];
                bool red = true;
                for (int q = c0 + lane; q < c1; q += 32)
                {
                    const int i = a.col_rows[q];
                    if (i < m0 && cover[i] < 2) red = false;
                }
                if (__all_sync(0xffffffffu, red))
                {
                    for (int q = c0 + lane; q < c1; q += 32)
                    {
                        const int i = a.col_rows[q];
                        if (i < m0) cover[i] -= 1;
                    }
                    if (lane == 0) state[j] = 0;
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }

    double acc = 0.0;
    int cnt = 0;
    for (int j = tid; j < n0; j += HT)
    {
        const unsigned char sel = (state[j] & ST_X) ? 1 : 0;
        a.cover_x[j] = sel;
        if (sel) { acc += a.c[j]; ++cnt; }
    }
    const double obj = block_sum(acc, sredd);
    const double nsel = block_sum((double)cnt, sredd);
    if (tid == 0)
    {
        const int ok = feasible && !s_bad;
        a.out->feasible = ok;
        a.out->cover_obj = ok ? obj : DBL_MAX;
        a.out->n_chosen = (int)nsel;
        a.out->repair_steps = s_bad ? -1 : s_steps;
        if (s_bad) { a.out->nif_feasible = 0; a.out->nif_obj = DBL_MAX; }
        if (a.host_flag)
        {
            __threadfence_system();
            *a.host_flag = a.flag_value;
        }
    }
}

__global__ void __launch_bounds__(HT, 1) k_node_heuristics_ref(HeurArgs a) { node_heuristics_ref_body(a); }

// a window of nodes in ONE launch: block b runs the node kernel on args[b] (same base model, own LP point and decisions)
__global__ void __launch_bounds__(HT, 1) k_node_heuristics_ref_batch(const HeurArgs *args)
{
    __shared__ HeurArgs sa;
    if (threadIdx.x == 0) sa = args[blockIdx.x];
    __syncthreads();
    node_heuristics_ref_body(sa);
}

} // namespace

size_t heur_smem_bytes(int m0, int n0, int rules)
{
    return (rules == SB200_HEUR_REFERENCE ? sizeof(double) * (size_t)m0 + (size_t)n0 : 0) +
           sizeof(int) * ((size_t)m0 + (size_t)n0) + (size_t)n0 + 16;
}

int launch_node_heuristics(const HeurArgs &a, cudaStream_t st)
{
    const size_t smem = heur_smem_bytes(a.m0, a.n0, a.rules);
    if (smem > 200 * 1024) return SB200_ERR_UNSUPPORTED;
    auto kern = a.rules == SB200_HEUR_REFERENCE ? k_node_heuristics_ref : k_node_heuristics;
    if (smem > 48 * 1024 &&      // per device, so not cached
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return SB200_ERR_CUDA;
    kern<<<1, HT, smem, st>>>(a);
    ++g_launch_count;
    return SB200_OK;
}

int launch_node_heuristics_batch(const HeurArgs *d_args, int count, int m0, int n0, cudaStream_t st)
{
    const size_t smem = heur_smem_bytes(m0, n0, SB200_HEUR_REFERENCE);
    if (smem > 200 * 1024) return SB200_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_node_heuristics_ref_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return SB200_ERR_CUDA;
    k_node_heuristics_ref_batch<<<count, HT, smem, st>>>(d_args);
    ++g_launch_count;
    return SB200_OK;
}

} // namespace sb200
