// oracle/refbuild/api_lp.cpp - TEST INFRASTRUCTURE (ours, not reference code).
//
// One LP relaxation (or a full branch-and-bound run) of an OR-Library set-covering file through the
// reference's PUBLIC model API, sypha::Solver (include/sypha/sypha.h:114-150), built the way
// examples/scp_solver.cpp builds it, but with the solver parameters taken from the command line and the
// results printed with 12 significant digits so that parity tests can compare the reference build and the
// drop-in build of the same tree.
//   api_lp <scp_file> [--lp] [--max-iter N] [--time-limit S] [--strategy auto|dense|krylov|...] [--verbosity V]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sypha/sypha.h"

int main(int argc, char **argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: %s <scp_file> [--lp] [--max-iter N] [--time-limit S] [--strategy S] [--verbosity V]\n", argv[0]);
        return 2;
    }
    bool lp = false;
    int maxIter = 100, verbosity = 0;
    double timeLimit = 0.0;
    std::string strategy = "auto";
    for (int i = 2; i < argc; ++i)
    {
        if (!std::strcmp(argv[i], "--lp")) lp = true;
        else if (!std::strcmp(argv[i], "--max-iter") && i + 1 < argc) maxIter = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--time-limit") && i + 1 < argc) timeLimit = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--strategy") && i + 1 < argc) strategy = argv[++i];
        else if (!std::strcmp(argv[i], "--verbosity") && i + 1 < argc) verbosity = std::atoi(argv[++i]);
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    FILE *fp = std::fopen(argv[1], "r");
    if (!fp) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    int m = 0, n = 0;
    if (std::fscanf(fp, "%d %d", &m, &n) != 2) return 1;
    std::vector<double> cost(n);
    for (int j = 0; j < n; ++j)
        if (std::fscanf(fp, "%lf", &cost[j]) != 1) return 1;

    sypha::Solver solver("SCP");
    sypha::SolverParameters &p = solver.parameters();
    p.verbosity = verbosity;
    p.mehrotra_max_iter = maxIter;
    p.disable_bnb = lp;
    p.bnb_hard_time_limit_sec = timeLimit;
    p.linear_solver_strategy = strategy;

    std::vector<sypha::Variable *> x(n);
    for (int j = 0; j < n; ++j) x[j] = solver.MakeBoolVar("x" + std::to_string(j));
    for (int i = 0; i < m; ++i)
    {
        int k = 0;
        if (std::fscanf(fp, "%d", &k) != 1) return 1;
        sypha::Constraint *row = solver.MakeRowConstraint(1.0, sypha::Solver::infinity(), "r" + std::to_string(i));
        for (int t = 0; t < k; ++t)
        {
            int j = 0;
            if (std::fscanf(fp, "%d", &j) != 1) return 1;
            row->SetCoefficient(x[j - 1], 1.0);
        }
    }
    std::fclose(fp);
    sypha::Objective *obj = solver.MutableObjective();
    obj->SetMinimization();
    for (int j = 0; j < n; ++j) obj->SetCoefficient(x[j], cost[j]);

    const sypha::ResultStatus st = solver.Solve();
    int selected = 0;
    double rounded = 0.0;
    for (int j = 0; j < n; ++j)
        if (x[j]->solution_value() > 0.5) { ++selected; rounded += cost[j]; }
    std::printf("{\"status\": %d, \"objective\": %.12g, \"dual_bound\": %.12g, \"iterations\": %d, \"wall_s\": %.6f, "
                "\"selected\": %d, \"selected_cost\": %.12g, \"m\": %d, \"n\": %d}\n",
                static_cast<int>(st), solver.objective_value(), solver.dual_objective_value(), solver.iterations(),
                solver.wall_time(), selected, rounded, m, n);
    return 0;
}
