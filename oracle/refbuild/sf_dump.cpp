// oracle/refbuild/sf_dump.cpp - TEST INFRASTRUCTURE (ours, not reference code).
//
// Prints the standard form the reference's OWN sypha::SolverImpl::buildStandardForm (src/sypha_api.cpp:136-250) builds
// for a model given on stdin, so that oracle/standard_form.py and sb200_build_standard_form can be pinned to the
// reference's code and not to a reading of it.  The reference's translation unit is compiled in place (included below,
// never copied); the model goes in through the public API (MakeNumVar / MakeRowConstraint / SetCoefficient / objective).
// Host code only: no CUDA call is made.
//   stdin:  n_vars n_rows maximize
//           n_obj  (var coef)*n_obj
//           per row: lb ub k (var coef)*k          ("inf" / "-inf" for absent bounds)
//   stdout: one JSON object {nrows, ncols, ncols_original, offs, inds, vals, obj, rhs}
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#define private public                      // Solver::impl_ (include/sypha/sypha.h:149)
#include "sypha_api.cpp"                    // -I$(REF)/src
#undef private

static double bound(const std::string &t)
{
    if (t == "inf" || t == "+inf") return std::numeric_limits<double>::infinity();
    if (t == "-inf") return -std::numeric_limits<double>::infinity();
    return std::atof(t.c_str());
}

template <class T> static void dump(const char *name, const std::vector<T> &v, bool last = false)
{
    std::printf("\"%s\": [", name);
    for (size_t i = 0; i < v.size(); ++i) std::printf(i ? ", %.17g" : "%.17g", (double)v[i]);
    std::printf(last ? "]" : "], ");
}

int main()
{
    int n_vars = 0, n_rows = 0, maximize = 0, n_obj = 0;
    if (!(std::cin >> n_vars >> n_rows >> maximize >> n_obj)) return 2;
    const auto t0 = std::chrono::steady_clock::now();
    sypha::Solver solver("sf");
    std::vector<sypha::Variable *> x((size_t)n_vars);
    for (int j = 0; j < n_vars; ++j) x[(size_t)j] = solver.MakeNumVar(0.0, sypha::Solver::infinity(), "x" + std::to_string(j));
    sypha::Objective *obj = solver.MutableObjective();
    for (int t = 0; t < n_obj; ++t)
    {
        int var = 0;
        double coef = 0.0;
        if (!(std::cin >> var >> coef)) return 2;
        obj->SetCoefficient(x[(size_t)var], coef);
    }
    if (maximize) obj->SetMaximization(); else obj->SetMinimization();
    for (int i = 0; i < n_rows; ++i)
    {
        std::string lb, ub;
        int k = 0;
        if (!(std::cin >> lb >> ub >> k)) return 2;
        sypha::Constraint *row = solver.MakeRowConstraint(bound(lb), bound(ub), "r" + std::to_string(i));
        for (int t = 0; t < k; ++t)
        {
            int var = 0;
            double coef = 0.0;
            if (!(std::cin >> var >> coef)) return 2;
            row->SetCoefficient(x[(size_t)var], coef);
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    const auto sf = solver.impl_->buildStandardForm();
    const auto t2 = std::chrono::steady_clock::now();
    if (std::getenv("SF_TIME"))
    {   // where the public API spends its time for a model of this size (input parsing included in the first figure)
        std::fprintf(stderr, "model through MakeNumVar / MakeRowConstraint / SetCoefficient: %.1f ms, buildStandardForm: %.1f ms\n",
                     std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
        return 0;
    }
    std::printf("{\"nrows\": %d, \"ncols\": %d, \"ncols_original\": %d, ", sf.nrows, sf.ncols, sf.ncolsOriginal);
    dump("offs", sf.csrOffs);
    dump("inds", sf.csrInds);
    dump("vals", sf.csrVals);
    dump("obj", sf.obj);
    dump("rhs", sf.rhs, true);
    std::printf("}\n");
    return 0;
}
