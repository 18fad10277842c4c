// oracle/refbuild/bnb_batched_main.cpp - TEST INFRASTRUCTURE (ours): runs the batched node loop
// (integration/sypha_bnb_batched_b200.cpp) on an OR-Library file the way the reference's own main does - a
// SyphaEnvironment, a SyphaNodeSparse, the reference's reader - and prints one JSON line.
//   bnb_batched <scp_file> [--slots K] [--max-iter N] [--max-nodes N] [--time-limit S] [--converged] [--verbosity V]
// SyphaEnvironment's parameters are private; the reference grants access to sypha::SolverImpl (sypha_environment.h:126),
// the class behind its public API.  This program does not link that API (sypha_api.cpp), so it defines the class itself
// to fill the same fields sypha_api.cpp:256-297 fills.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#include "sypha_environment.h"
#include "sypha_logger.h"
#include "sypha_node_sparse.h"
#include "sypha_bnb_batched_b200.h"

namespace sypha
{
class SolverImpl
{
public:
    static SyphaStatus configure(SyphaEnvironment &env, const std::string &path, int verbosity, int maxIter, double timeLimit)
    {
        env.setDefaultParameters();
        env.verbosityLevel = verbosity;
        env.mehrotraMaxIter = maxIter;
        env.bnbHardTimeLimitSeconds = timeLimit;
        env.modelType = MODEL_TYPE_SCP;
        env.inputFilePath = path;
        env.sparse = true;
        env.internalStatus = CODE_SUCCESSFUL;
        env.logger_ = std::make_unique<SyphaLogger>(env.timer(), verbosity <= 0 ? LOG_ERROR : (verbosity <= 5 ? LOG_INFO : LOG_DEBUG));
        if (timeLimit > 0.0) env.logger_->setHardTimeLimit(timeLimit * 1000.0);
        return env.setUpDevice();
    }
};
} // namespace sypha

int main(int argc, char **argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <scp_file> [--slots K] [--max-iter N] [--max-nodes N] [--time-limit S] [--converged]\n", argv[0]); return 2; }
    SyphaBatchedBnbConfig cfg;
    int verbosity = 0, maxIter = 100;
    double timeLimit = 0.0;
    for (int i = 2; i < argc; ++i)
    {
        if (!std::strcmp(argv[i], "--slots") && i + 1 < argc) cfg.slots = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--max-iter") && i + 1 < argc) maxIter = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--max-nodes") && i + 1 < argc) cfg.maxNodes = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--time-limit") && i + 1 < argc) timeLimit = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--verbosity") && i + 1 < argc) verbosity = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--converged")) cfg.nodeLpToConvergence = true;
        else if (!std::strcmp(argv[i], "--continuous")) cfg.continuousBatching = true;
        else if (!std::strcmp(argv[i], "--windows-in-flight") && i + 1 < argc) cfg.windowsInFlight = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--no-preprocessing")) cfg.referencePreprocessing = false;
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    SyphaEnvironment env;
    if (sypha::SolverImpl::configure(env, argv[1], verbosity, maxIter, timeLimit) != CODE_SUCCESSFUL) return 1;
    SyphaNodeSparse node(env);
    if (node.readModel() != CODE_SUCCESSFUL) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    const int mIn = node.nrows, nIn = node.ncolsOriginal;
    SyphaBatchedBnbStats st;
    if (solver_sparse_branch_and_bound_batched(node, cfg, &st) != CODE_SUCCESSFUL) return 1;
    int selected = 0;
    for (double v : node.hX) selected += v > 0.5 ? 1 : 0;
    std::printf("{\"objective\": %.12g, \"dual_bound\": %.12g, \"mip_gap\": %.6g, \"nodes\": %d, \"lp_iterations\": %d, "
                "\"nodes_per_sec\": %.2f, \"wall_ms\": %.2f, \"lp_device_ms_per_node\": %.3f, \"open_nodes\": %d, \"pruned\": %d, "
                "\"failed_lps\": %d, \"integral_nodes\": %d, \"dropped_too_deep\": %d, \"greedy_incumbent\": %.12g, \"root_bound\": %.9g, "
                "\"selected\": %d, \"m\": %d, \"n\": %d, \"base_cols\": %d, \"slots\": %d}\n",
                node.objvalPrim, node.objvalDual, node.mipGap, st.processedNodes, st.totalLpIterations, st.nodesPerSecond, st.wallMs,
                st.processedNodes ? st.lpDeviceMs / st.processedNodes : 0.0, st.openNodes, st.prunedByBound, st.failedLps,
                st.integralNodes, st.droppedTooDeep, st.greedyIncumbent, st.rootBound, selected, mIn, nIn, st.baseColsOriginal, cfg.slots);
    return 0;
}
