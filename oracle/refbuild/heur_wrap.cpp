// oracle/refbuild/heur_wrap.cpp - TEST INFRASTRUCTURE (ours): a C entry point around the reference's OWN
// integer heuristics and branching selectors, so that tests can call the unmodified reference code
// (src/sypha_solver_heuristics.cpp, compiled where it lies by oracle/Makefile into oracle/_ref/libref_heur.so)
// on arrays.  The fractional-candidate list is built as collect_fractional_candidates does
// (src/sypha_solver_bnb.cpp:368-382; that translation unit drags in CUDA, so its nine lines are repeated here).
#include <cmath>
#include <string>
#include <vector>

#include "sypha_solver_heuristics.h"

extern "C" {

// which: 0 = nearest_integer_fixing, 1 = dual_guided_cover_repair.  Returns feasible (0/1).
int ref_heuristic(int which, int m, int n, int n0, const int *offs, const int *inds, const double *vals,
                  const double *obj, const double *rhs, const double *x, int nx, const double *y, int ny, int ndec,
                  const int *dvar, const int *dfix, double tol, double *sol, double *objective)
{
    BaseRelaxationModel base;
    base.nrows = m;
    base.ncols = n;
    base.ncolsOriginal = n0;
    base.ncolsInputOriginal = n0;
    base.nnz = offs[m];
    base.csrOffs.assign(offs, offs + m + 1);
    base.csrInds.assign(inds, inds + offs[m]);
    base.csrVals.assign(vals, vals + offs[m]);
    base.obj.assign(obj, obj + n);
    base.rhs.assign(rhs, rhs + m);
    BranchNodeState node;
    for (int d = 0; d < ndec; ++d) node.decisions.push_back(BranchDecision{dvar[d], dfix[d]});
    std::vector<double> px(x, x + nx), py(y, y + ny);
    auto hs = makeIntegerHeuristics(which == 0 ? "nearest_integer_fixing" : "dual_guided_cover_repair");
    IntegerHeuristicResult r = hs[0]->tryBuild(px, py, base, node, tol);
    for (int j = 0; j < n0; ++j) sol[j] = j < (int)r.solution.size() ? r.solution[(size_t)j] : 0.0;
    *objective = r.objective;
    return r.feasible ? 1 : 0;
}

// rule: 0 = most_fractional, 1 = highest_cost_fractional.  Returns the column or -1.
int ref_select_branch(int rule, const double *x, const double *obj, int n0, double tol)
{
    std::vector<double> px(x, x + n0), pobj(obj, obj + n0);
    std::vector<int> cand;
    for (int j = 0; j < n0; ++j)
    {
        const double v = px[(size_t)j];
        const double nearest = floor(v + 0.5);
        if ((fabs(v - nearest) > tol) || (nearest < -tol) || (nearest > 1.0 + tol)) cand.push_back(j);
    }
    if (cand.empty()) return -1;
    return makeBranchSelector(rule == 0 ? "most_fractional" : "highest_cost_fractional")->select(px, pobj, cand);
}
}
