// sypha_bnb_batched_b200.cpp - reference-side C++: the branch-and-bound node loop of
// /root/reference/src/sypha_solver_bnb_driver.cpp:698-1046 restructured so that K node LPs are in flight at once.
//
// The reference pops ONE node, rebuilds its CSR on the host (build_branch_model, sypha_solver_bnb.cpp:418-490), re-uploads it
// (copyModelOnDevice: 5 frees, 5 mallocs, the whole model) and solves its LP before it looks at the next node.  Here the
// driver's search logic is kept - FIFO frontier, pruning by the parent's bound, the reliability rule for a node's bound
// (bnb_driver.cpp:866-877), bound tightening for integral costs, the two integer heuristics in their order, the configured
// branching selector, append_decision_if_consistent - and only the node BODY changes: the base model stays resident in K
// workspaces of libsypha_b200, a node travels as its decision list (sb200_node_delta, 20 bytes per decision), its LP is one
// launch of one thread block (the throughput form, sb200_set_concurrency_hint), the branching variable and both heuristics
// run on the device behind it (sb200_node_heuristics, the reference's rules bit for bit), and sb200_solve_stream hands a
// slot its next node the moment it is free.  Everything else is the reference's own code, called unchanged:
// greedy_set_cover_heuristic, SyphaNodeSparse::reduceByIncumbent / applyIncumbentBudgetPruning / applyCostDrivenReduction /
// applyDominancePreprocessing, has_integer_objective, tighten_dual_bound, compute_mip_gap, append_decision_if_consistent.
//
// Not carried over (documented in INTEGRATION.md): root cut rounds (cut rows carry coefficients > 1: the node LPs would
// leave the unit-coefficient fast path), mid-search column removal, and the adaptive iteration cap.
//
// Build: oracle/Makefile links it with the unmodified reference objects into oracle/_ref/bnb_batched_b200.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <limits>
#include <vector>

#include "sypha_bnb_batched_b200.h"

#include "sypha_environment.h"
#include "sypha_node_sparse.h"
#include "sypha_preprocessor.h"
#include "sypha_solver_bnb.h"
#include "sypha_solver_heuristics.h"
#include "sypha_solver_sparse.h"

#include "sypha_b200.h"

namespace
{
struct Search
{
    SyphaNodeSparse *node = nullptr;
    SyphaLogger *log = nullptr;
    const SyphaBatchedBnbConfig *cfg = nullptr;
    SyphaBatchedBnbStats st;
    BaseRelaxationModel base;
    std::deque<BranchNodeState> frontier;
    std::vector<BranchNodeState> inSlot;
    std::vector<std::vector<int>> slotVar;
    std::vector<std::vector<double>> slotCoef, slotRhs;
    std::vector<sb200_ws *> ws;
    double bestObj = std::numeric_limits<double>::infinity();
    std::vector<double> bestSolution;       // input-original column space
    double tol = 1e-12, intTol = 1e-6;
    bool objIsIntegral = false;
    int maxNodes = 0, maxDepth = 64;
    double deadlineMs = 0.0;
    bool stopped = false;
    std::vector<unsigned char> cover;
};

void adoptIncumbent(Search &S, double obj, const unsigned char *x01)
{   // active column space -> input-original space (adoptIncumbentSolution, bnb_driver.cpp:20-40)
    S.bestObj = obj;
    S.bestSolution.assign(static_cast<size_t>(S.base.ncolsInputOriginal), 0.0);
    for (int j = 0; j < S.base.ncolsOriginal; ++j)
        if (x01[j])
        {
            const int in = S.base.activeToOriginalCol[static_cast<size_t>(j)];
            if (in >= 0 && in < S.base.ncolsInputOriginal) S.bestSolution[static_cast<size_t>(in)] = 1.0;
        }
    if (S.log) S.log->log(LOG_INFO, "New incumbent found: %.12g", obj);
}

int nextNode(void *user, int slot, sb200_node_delta *delta)
{
    Search &S = *static_cast<Search *>(user);
    if (S.stopped) return 0;
    if (S.st.processedNodes + S.st.inFlight >= S.maxNodes ||
        (S.deadlineMs > 0.0 && S.node->env->timer() >= S.deadlineMs) ||
        (S.log && S.log->isStopRequested()))
    {
        S.stopped = true;
        return 0;
    }
    while (!S.frontier.empty())
    {
        BranchNodeState nd = std::move(S.frontier.front());
        S.frontier.pop_front();                                     // FIFO (sypha_solver_bnb.cpp:42-43)
        if (nd.parentDualBound >= S.bestObj - S.tol)                // bnb_driver.cpp:797
        {
            ++S.st.prunedByBound;
            continue;
        }
        const int k = static_cast<int>(nd.decisions.size());
        if (k > S.maxDepth)
        {
            ++S.st.droppedTooDeep;                                  // reported; raise maxDepth for deeper searches
            continue;
        }
        std::vector<int> &var = S.slotVar[static_cast<size_t>(slot)];
        std::vector<double> &coef = S.slotCoef[static_cast<size_t>(slot)], &rhs = S.slotRhs[static_cast<size_t>(slot)];
        var.resize(static_cast<size_t>(k));
        coef.resize(static_cast<size_t>(k));
        rhs.resize(static_cast<size_t>(k));
        for (int r = 0; r < k; ++r)
        {   // build_branch_model's row: (fix == 0 ? -1 : +1) x_var - slack = fix   (sypha_solver_bnb.cpp:453-468)
            var[static_cast<size_t>(r)] = nd.decisions[static_cast<size_t>(r)].varIndex;
            coef[static_cast<size_t>(r)] = nd.decisions[static_cast<size_t>(r)].fixValue == 0 ? -1.0 : 1.0;
            rhs[static_cast<size_t>(r)] = static_cast<double>(nd.decisions[static_cast<size_t>(r)].fixValue);
        }
        delta->n_extra_rows = k;
        delta->var = var.data();
        delta->coef = coef.data();
        delta->rhs = rhs.data();
        S.inSlot[static_cast<size_t>(slot)] = std::move(nd);
        ++S.st.inFlight;
        return 1;
    }
    return 0;
}

void nodeDone(void *user, int slot, const sb200_result *r, const sb200_heur_result *h)
{
    Search &S = *static_cast<Search *>(user);
    const BranchNodeState &nd = S.inSlot[static_cast<size_t>(slot)];
    --S.st.inFlight;
    if (r->status != SB200_OK)
    {   // a failed LP below the root is skipped (bnb_driver.cpp:844-859)
        ++S.st.failedLps;
        return;
    }
    ++S.st.processedNodes;
    S.st.totalLpIterations += r->iterations;
    S.st.lpDeviceMs += r->ms_start + r->ms_setup + r->ms_loop;
    const bool consistent = std::isfinite(r->dual_obj) && std::isfinite(r->primal_obj) && r->dual_obj <= r->primal_obj + S.tol;
    const bool reliable = r->reason == SB200_TERM_CONVERGED && consistent;                         // :866-870
    double bound = reliable ? r->dual_obj : nd.parentDualBound, boundRaw = reliable ? r->dual_obj : nd.parentDualBoundRaw;
    if (S.objIsIntegral && reliable && std::isfinite(bound)) bound = tighten_dual_bound(bound, S.intTol);
    if (nd.decisions.empty()) S.st.rootBound = boundRaw;
    // heuristics in the configured order; the first that improves the incumbent is taken (:885-903)
    if (h->nif_feasible && h->nif_obj < S.bestObj - S.tol)
    {
        sb200_get_rounded(S.ws[static_cast<size_t>(slot)], S.cover.data());
        adoptIncumbent(S, h->nif_obj, S.cover.data());
    }
    else if (h->feasible && h->cover_obj < S.bestObj - S.tol)
    {
        sb200_get_cover(S.ws[static_cast<size_t>(slot)], S.cover.data());
        adoptIncumbent(S, h->cover_obj, S.cover.data());
    }
    if (bound >= S.bestObj - S.tol)
    {
        ++S.st.prunedByBound;
        return;
    }
    if (h->branch_var < 0)
    {   // integral LP point (is_binary_integral_solution): an incumbent candidate at its own cost
        ++S.st.integralNodes;
        if (h->rounded_obj < S.bestObj - S.tol)
        {
            sb200_get_rounded(S.ws[static_cast<size_t>(slot)], S.cover.data());
            adoptIncumbent(S, h->rounded_obj, S.cover.data());
        }
        return;
    }
    for (int value = 0; value <= 1; ++value)
    {
        BranchNodeState child;
        if (append_decision_if_consistent(nd, h->branch_var, value, &child))
        {
            child.parentDualBound = bound;
            child.parentDualBoundRaw = boundRaw;
            S.frontier.push_back(std::move(child));
        }
    }
}

[[noreturn]] void fatal(const char *what, int code, sb200_ws *h)
{
    fprintf(stderr, "sypha_b200: %s failed (code %d): %s\n", what, code, h ? sb200_last_error(h) : "");
    exit(EXIT_FAILURE);
}
} // namespace

SyphaStatus solver_sparse_branch_and_bound_batched(SyphaNodeSparse &node, const SyphaBatchedBnbConfig &cfg,
                                                   SyphaBatchedBnbStats *statsOut)
{
    Search S;
    S.node = &node;
    S.cfg = &cfg;
    S.log = node.env->getLogger();
    S.tol = node.env->getPxTolerance();
    S.intTol = node.env->getBnbIntegralityTol();
    S.maxDepth = cfg.maxDepth > 0 ? cfg.maxDepth : 64;
    S.maxNodes = cfg.maxNodes > 0 ? cfg.maxNodes : node.env->getBnbMaxNodes();
    if (node.ncolsInputOriginal <= 0) node.ncolsInputOriginal = node.ncolsOriginal;
    const int ncolsInput = node.ncolsInputOriginal;

    // ---- the reference's prelude, unchanged code (bnb_driver.cpp:262-334) ------------------------------------------------
    GreedySetCoverResult greedy = greedy_set_cover_heuristic(node.nrows, node.ncolsOriginal, node.hCsrMatInds, node.hCsrMatOffs,
                                                             node.hCsrMatVals, node.hObjDns.data());
    if (greedy.feasible)
    {
        S.bestObj = greedy.objective;
        S.bestSolution.assign(static_cast<size_t>(ncolsInput), 0.0);
        for (int col : greedy.selectedColumns)
        {
            const int in = node.hActiveToInputCols.empty() ? col : node.hActiveToInputCols[static_cast<size_t>(col)];
            if (in >= 0 && in < ncolsInput) S.bestSolution[static_cast<size_t>(in)] = 1.0;
        }
        S.st.greedyIncumbent = greedy.objective;
        if (S.log) S.log->log(LOG_INFO, "Greedy heuristic incumbent: %.12g", S.bestObj);
        node.reduceByIncumbent(S.bestObj);
        node.applyIncumbentBudgetPruning(S.bestObj);
    }
    if (cfg.referencePreprocessing)
    {
        node.applyCostDrivenReduction();
        node.applyDominancePreprocessing();
    }
    S.objIsIntegral = has_integer_objective(node.hObjDns.data(), node.ncolsOriginal, S.intTol);

    // ---- base model (what buildBaseModel copies, bnb_driver.cpp:166-190) ---------------------------------------------------
    BaseRelaxationModel &base = S.base;
    base.nrows = node.nrows;
    base.ncols = node.ncols;
    base.ncolsOriginal = node.ncolsOriginal;
    base.ncolsInputOriginal = ncolsInput;
    base.nnz = node.nnz;
    base.csrInds = node.hCsrMatInds;
    base.csrOffs = node.hCsrMatOffs;
    base.csrVals = node.hCsrMatVals;
    base.obj = node.hObjDns;
    base.rhs = node.hRhsDns;
    base.activeToOriginalCol = node.hActiveToInputCols;
    if (base.activeToOriginalCol.empty())
    {
        base.activeToOriginalCol.resize(static_cast<size_t>(base.ncolsOriginal));
        for (int j = 0; j < base.ncolsOriginal; ++j) base.activeToOriginalCol[static_cast<size_t>(j)] = j;
    }
    S.st.baseRows = base.nrows;
    S.st.baseColsOriginal = base.ncolsOriginal;
    S.cover.assign(static_cast<size_t>(base.ncolsOriginal) + 1, 0);

    // ---- K resident copies of the base model ---------------------------------------------------------------------------
    const int K = std::max(1, cfg.slots);
    int dev = 0;
    cudaGetDevice(&dev);
    sb200_caps caps;
    caps.m_max = base.nrows + S.maxDepth;
    caps.n_max = base.ncols + S.maxDepth;
    caps.nnz_max = static_cast<long long>(base.nnz) + 2LL * S.maxDepth;
    // windows in flight: P sets of K slots used alternately (slot index = set * K + position)
    const int P = (cfg.continuousBatching || K < 2) ? 1 : std::max(1, cfg.windowsInFlight);
    const int KP = K * P;
    S.ws.resize(static_cast<size_t>(KP), nullptr);
    S.inSlot.resize(static_cast<size_t>(KP));
    S.slotVar.resize(static_cast<size_t>(KP));
    S.slotCoef.resize(static_cast<size_t>(KP));
    S.slotRhs.resize(static_cast<size_t>(KP));
    const int branchRule = node.env->getBnbVarSelectionStrategy() == "highest_cost_fractional"
                               ? SB200_BRANCH_HIGHEST_COST_FRACTIONAL : SB200_BRANCH_MOST_FRACTIONAL;
    for (int i = 0; i < KP; ++i)
    {
        int rc = sb200_ws_create(dev, &caps, &S.ws[static_cast<size_t>(i)]);
        if (rc != SB200_OK) fatal("sb200_ws_create", rc, nullptr);
        sb200_ws *w = S.ws[static_cast<size_t>(i)];
        rc = sb200_load_model(w, base.nrows, base.ncols, base.ncolsOriginal, base.nnz, base.csrOffs.data(), base.csrInds.data(),
                              base.csrVals.data(), base.obj.data(), base.rhs.data(), /*ptrs_on_device=*/0, SB200_STRATEGY_CHOLESKY);
        if (rc != SB200_OK) fatal("sb200_load_model", rc, w);
        sb200_set_concurrency_hint(w, K);
        sb200_set_heuristic_rules(w, SB200_HEUR_REFERENCE, branchRule, S.intTol);
        rc = sb200_prepare_nodes(w, S.maxDepth);           // no allocation inside the search (it would wait for a window in flight)
        if (rc != SB200_OK) fatal("sb200_prepare_nodes", rc, w);
    }

    // ---- node LP configuration: the reference's (bnb_driver.cpp:833-837) ---------------------------------------------------------
    sb200_params p;
    sb200_default_params(&p);
    p.max_iter = cfg.maxIterations > 0 ? cfg.maxIterations : node.env->getMehrotraMaxIter();
    p.eta = node.env->getMehrotraEta();
    p.mu_tol = node.env->getMehrotraMuTol();
    p.gap_enabled = cfg.nodeLpToConvergence ? 0 : 1;
    p.gap_window = node.env->getBnbGapStallBranchIters();
    p.gap_min_improv_pct = node.env->getBnbGapStallMinImprovPct();

    BranchNodeState root;
    S.frontier.push_back(root);
    node.timeSolverStart = node.env->timer();
    const double limit = node.env->getBnbHardTimeLimitSeconds();
    S.deadlineMs = limit > 0.0 ? node.timeSolverStart + 1000.0 * limit : 0.0;
    if (S.log) S.log->log(LOG_INFO, "Branch-and-bound started (%d node LPs in flight)", K);
    const auto t0 = std::chrono::steady_clock::now();
    int rc = SB200_OK;
    if (cfg.continuousBatching)
        rc = sb200_solve_stream(S.ws.data(), K, &p, nextNode, nodeDone, &S);
    else
    {   // windows of up to K nodes (the reference's DeviceNodeWindow, sypha_solver_bnb.cpp:32-69, pops one at a time):
        // K one-block LPs launched together, the K node kernels behind them, then the host-side node rule.  With P > 1
        // sets of slots the next window is already on the GPU while this loop branches on the finished one.
        std::vector<sb200_node_delta> deltas(static_cast<size_t>(KP));
        std::vector<sb200_result> results(static_cast<size_t>(KP));
        std::vector<sb200_heur_result> heur(static_cast<size_t>(KP));
        struct Window { int set, cnt; };
        const bool traceWindows = std::getenv("SB200_TRACE_WINDOWS") != nullptr;
        std::deque<Window> inFlight;
        std::vector<int> freeSets;
        for (int s = P - 1; s >= 0; --s) freeSets.push_back(s);
        while (rc == SB200_OK)
        {
            while (rc == SB200_OK && !freeSets.empty())
            {
                const int set = freeSets.back(), base = set * K;
                int cnt = 0;
                while (cnt < K)
                {
                    deltas[static_cast<size_t>(base + cnt)] = sb200_node_delta{};
                    if (!nextNode(&S, base + cnt, &deltas[static_cast<size_t>(base + cnt)])) break;
                    results[static_cast<size_t>(base + cnt)] = sb200_result{};
                    ++cnt;
                }
                if (cnt == 0) break;
                sb200_ws **w = S.ws.data() + base;
                rc = sb200_window_begin(w, cnt, deltas.data() + base, &p, results.data() + base, 1);
                if (rc == SB200_ERR_UNSUPPORTED)
                {   // not a one-launch window (one node, or a node deeper than a thread block takes): the deltas are applied,
                    // solve it here and now
                    rc = sb200_solve_batch(w, cnt, nullptr, &p, results.data() + base);
                    if (rc == SB200_OK) rc = sb200_node_heuristics(w, cnt, heur.data() + base);
                    if (rc != SB200_OK) break;
                    for (int i = 0; i < cnt; ++i)
                        nodeDone(&S, base + i, &results[static_cast<size_t>(base + i)], &heur[static_cast<size_t>(base + i)]);
                    continue;
                }
                if (rc != SB200_OK) break;
                freeSets.pop_back();
                inFlight.push_back(Window{set, cnt});
            }
            if (rc != SB200_OK || inFlight.empty()) break;
            const Window win = inFlight.front();
            inFlight.pop_front();
            const int base = win.set * K;
            const auto tf0 = std::chrono::steady_clock::now();
            rc = sb200_window_finish(S.ws.data() + base, win.cnt, results.data() + base, heur.data() + base);
            if (rc != SB200_OK) break;
            const auto tf1 = std::chrono::steady_clock::now();
            for (int i = 0; i < win.cnt; ++i)
                nodeDone(&S, base + i, &results[static_cast<size_t>(base + i)], &heur[static_cast<size_t>(base + i)]);
            freeSets.push_back(win.set);
            if (traceWindows)
            {
                double wms = 0.0;
                int wl = 0;
                sb200_last_window(S.ws[static_cast<size_t>(base)], &wms, &wl);
                std::fprintf(stderr, "[window] set %d, %d nodes: waited %.2f ms, kernel %.2f ms, node rule %.2f ms, at %.1f ms\n", win.set, win.cnt,
                             std::chrono::duration<double, std::milli>(tf1 - tf0).count(), wms,
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tf1).count(),
                             std::chrono::duration<double, std::milli>(tf1 - t0).count());
            }
        }
        while (!inFlight.empty())
        {   // an error path: let the GPU finish what was launched before the workspaces go away
            const Window win = inFlight.front();
            inFlight.pop_front();
            sb200_window_finish(S.ws.data() + win.set * K, win.cnt, results.data() + win.set * K, nullptr);
        }
    }
    const double wallMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc != SB200_OK) fatal("sb200_solve_stream", rc, S.ws[0]);
    node.timeSolverEnd = node.env->timer();

    // ---- results, as the reference leaves them (bnb_driver.cpp:1076-1110) ------------------------------------------------------
    double lower = std::numeric_limits<double>::infinity();
    for (const BranchNodeState &nd : S.frontier) lower = std::min(lower, nd.parentDualBound);
    node.iterations = S.st.totalLpIterations;
    if (std::isfinite(S.bestObj))
    {
        node.objvalPrim = S.bestObj;
        node.hX = S.bestSolution;
    }
    else
        node.objvalPrim = std::numeric_limits<double>::infinity();
    const bool exhausted = S.frontier.empty() && !S.stopped && S.st.droppedTooDeep == 0;
    if (std::isfinite(S.bestObj) && exhausted)
    {
        node.objvalDual = S.bestObj;
        node.mipGap = 0.0;
        if (S.log) S.log->log(LOG_INFO, "Optimality proven: search frontier exhausted");
    }
    else
    {
        node.objvalDual = std::isfinite(lower) ? lower : S.st.rootBound;
        node.mipGap = compute_mip_gap(node.objvalPrim, node.objvalDual);
    }
    S.st.wallMs = wallMs;
    S.st.nodesPerSecond = wallMs > 0.0 ? 1e3 * S.st.processedNodes / wallMs : 0.0;
    S.st.openNodes = static_cast<int>(S.frontier.size());
    S.st.incumbent = S.bestObj;
    if (S.log) S.log->log(LOG_INFO, "BnB processed %d nodes, %d total LP iterations", S.st.processedNodes, S.st.totalLpIterations);
    for (sb200_ws *w : S.ws) sb200_ws_destroy(w);
    if (statsOut) *statsOut = S.st;
    return CODE_SUCCESSFUL;
}
