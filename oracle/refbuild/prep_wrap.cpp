// oracle/refbuild/prep_wrap.cpp - TEST INFRASTRUCTURE (ours, not reference code): a C entry point around the reference's
// greedy_set_cover_heuristic (src/sypha_preprocessor.cpp:11-96, compiled in place by oracle/Makefile) so that the Python
// restatement the B&B bench uses for its first incumbent (sypha_b200/bnb.py::greedy_cover) can be pinned to it.
#include <vector>

#include "sypha_preprocessor.h"

extern "C" int ref_greedy_set_cover(int nrows, int ncols_original, const int *csr_offs, const int *csr_inds, const double *csr_vals,
                                    const double *obj, int *selected /* [ncols_original] */, int *n_selected, double *objective)
{
    const int nnz = csr_offs[nrows];
    const std::vector<int> inds(csr_inds, csr_inds + nnz), offs(csr_offs, csr_offs + nrows + 1);
    const std::vector<double> vals(csr_vals, csr_vals + nnz);
    const GreedySetCoverResult r = greedy_set_cover_heuristic(nrows, ncols_original, inds, offs, vals, obj);
    *n_selected = (int)r.selectedColumns.size();
    for (int i = 0; i < *n_selected; ++i) selected[i] = r.selectedColumns[(size_t)i];
    *objective = r.objective;
    return r.feasible ? 1 : 0;
}
