// sypha_solver_b200.cpp - reference-side shim: the unchanged entry points of the IPM hot path,
// implemented over the C ABI of libsypha_b200.so (include/sypha_b200.h).
//
// Drop-in replacement for these reference translation units (paths under /root/reference/src):
//     sypha_solver.cpp            solver_sparse_mehrotra, solver_sparse_mehrotra_run   (:25-886)
//     sypha_solver_workspace.cpp  initializeIpmWorkspace, releaseIpmWorkspace          (:5-89)
//     sypha_solver_init.cpp       solver_sparse_mehrotra_init_gsl (now on the GPU)     (:543-652)
//     sypha_solver_dense_linear.cpp, sypha_solver_krylov.cu, sypha_solver_utils.cu     (not needed)
// Callers stay as they are: sypha_api.cpp:346, sypha_node_sparse.cpp:139,
// sypha_solver_bnb_driver.cpp:352,474,626,843,1152,1161.
//
// Build:  g++ -std=c++17 -I<sypha>/src -I<this repo>/include \
//     -I<this repo>/integration/stubs -I$CUDA/include -c sypha_solver_b200.cpp ; link -lsypha_b200
//
// The reference's IpmWorkspace (sypha_solver.h:76-105) is kept byte for byte - the B&B driver holds
// one BY VALUE (bnb_driver.cpp:618).  Its `krylov` slot (an opaque pointer the callers never touch)
// carries the sb200_ws handle; `isAllocated` keeps its meaning.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include "sypha_solver.h"
#include "sypha_solver_sparse.h"
#include "sypha_node_sparse.h"
#include "sypha_environment.h"

#include "sypha_b200.h"

namespace
{
inline sb200_ws *handle_of(IpmWorkspace *ws) { return reinterpret_cast<sb200_ws *>(ws->krylov); }

// CUDA / library failures are fatal in the reference (checkCudaErrors -> exit, sypha_cuda_helper.h:19-31);
// the C ABI only returns codes, the shim keeps the reference's convention.
[[noreturn]] void fatal(const char *what, int code, sb200_ws *h)
{
    fprintf(stderr, "sypha_b200: %s failed (code %d): %s\n", what, code, h ? sb200_last_error(h) : "");
    exit(EXIT_FAILURE);
}

int strategy_of(const std::string &s)
{
    if (s == "cholesky" || s == "dense") return SB200_STRATEGY_CHOLESKY;
    if (s == "syrk") return SB200_STRATEGY_SYRK;
    if (s == "pcg" || s == "krylov") return SB200_STRATEGY_PCG;
    return SB200_STRATEGY_AUTO;       // "auto", "sparse_qr"
}
} // namespace

void initializeIpmWorkspace(IpmWorkspace *ws, int maxKktNrows, int maxKktNnz, int maxNcols)
{
    // sizing arguments are KKT-shaped (bnb_driver.cpp:620-626): rows 2n+m, nnz 2nnz+3n
    sb200_caps caps;
    caps.n_max = maxNcols;
    caps.m_max = maxKktNrows - 2 * maxNcols;
    caps.nnz_max = (static_cast<long long>(maxKktNnz) - 3LL * maxNcols) / 2;
    if (ws->isAllocated && ws->krylov) return;     // grow-only: sb200_load_model grows on demand
    int dev = 0;
    cudaGetDevice(&dev);
    sb200_ws *h = nullptr;
    const int rc = sb200_ws_create(dev, (caps.m_max > 0 && caps.nnz_max > 0) ? &caps : nullptr, &h);
    if (rc != SB200_OK) fatal("sb200_ws_create", rc, nullptr);
    ws->krylov = reinterpret_cast<KrylovSolveWorkspace *>(h);
    ws->kktNrowsCapacity = maxKktNrows;
    ws->kktNnzCapacity = maxKktNnz;
    ws->vectorCapacity = maxKktNrows;
    ws->isAllocated = true;
}

void releaseIpmWorkspace(IpmWorkspace *ws)
{
    if (ws->krylov) sb200_ws_destroy(handle_of(ws));
    *ws = IpmWorkspace();
}

SyphaStatus solver_sparse_mehrotra_run(SyphaNodeSparse &node, const SolverExecutionConfig &config,
                                       SolverExecutionResult *result, IpmWorkspace *workspace)
{
    const bool useWs = (workspace != nullptr) && workspace->isAllocated;     // sypha_solver.cpp:63
    IpmWorkspace local;
    IpmWorkspace *ws = workspace;
    if (!useWs)
    {
        ws = &local;
        initializeIpmWorkspace(ws, 2 * node.ncols + node.nrows, 2 * node.nnz + 3 * node.ncols, node.ncols);
    }
    sb200_ws *h = handle_of(ws);
    SyphaEnvironment *env = node.env;

    // model: the device CSR that copyModelOnDevice() uploaded (sypha_node_sparse.cpp:156-198)
    node.timePreSolStart = env->timer();
    int rc = sb200_load_model(h, node.nrows, node.ncols, node.ncolsOriginal, node.nnz, node.dCsrMatOffs,
                              node.dCsrMatInds, node.dCsrMatVals, node.dObjDns, node.dRhsDns,
                              /*ptrs_on_device=*/1, strategy_of(env->getLinearSolverStrategy()));
    if (rc != SB200_OK) fatal("sb200_load_model", rc, h);

    sb200_params p;
    sb200_default_params(&p);
    p.max_iter = config.maxIterations > 0 ? config.maxIterations : env->getMehrotraMaxIter();   // :488
    p.eta = env->getMehrotraEta();
    p.mu_tol = env->getMehrotraMuTol();
    p.gap_enabled = config.gapStagnation.enabled ? 1 : 0;
    p.gap_window = config.gapStagnation.windowIterations;
    p.gap_min_improv_pct = config.gapStagnation.minImprovementPct;
    p.cg_max_iter = env->getKrylovMaxCgIter();
    p.cg_tol_initial = env->getKrylovCgTolInitial();
    p.cg_tol_final = env->getKrylovCgTolFinal();
    p.cg_tol_decay = env->getKrylovCgTolDecayRate();

    // the logger's watchdog flag (sypha_solver.cpp:498-502) is an std::atomic<bool> behind a getter: the
    // library asks it through a callback every time it looks at the LP's scalar block (every iteration)
    if (env->getLogger())
    {
        p.stop_cb = [](void *lg) -> int { return static_cast<SyphaLogger *>(lg)->isStopRequested() ? 1 : 0; };
        p.stop_user = env->getLogger();
    }

    node.hX.resize(node.ncols);
    node.hY.resize(node.nrows);
    node.hS.resize(node.ncols);
    sb200_result r = {};
    if (result != nullptr)
    {
        result->primalSolution.resize(static_cast<size_t>(node.ncols), 0.0);
        result->dualSolution.resize(static_cast<size_t>(node.nrows), 0.0);
        r.x_host = result->primalSolution.data();
        r.y_host = result->dualSolution.data();
    }
    r.x0_host = node.hX.data();      // starting point, sypha_solver.cpp:72-78
    r.y0_host = node.hY.data();
    r.s0_host = node.hS.data();

    rc = sb200_solve(h, &p, &r);
    if (rc != SB200_OK) fatal("sb200_solve", rc, h);

    // node.* outputs, sypha_solver.cpp:774-821
    const double t1 = env->timer();
    node.timeSolverEnd = t1;
    node.timeSolverStart = t1 - r.ms_loop;
    node.timePreSolEnd = node.timeSolverStart;
    node.timeStartSolEnd = node.timePreSolEnd - r.ms_setup;
    node.timeStartSolStart = node.timeStartSolEnd - r.ms_start;
    node.iterations = r.iterations;
    node.objvalPrim = r.primal_obj;
    node.objvalDual = r.dual_obj;
    node.mipGap = std::numeric_limits<double>::infinity();
    const bool numerical = (r.status != SB200_OK);
    if (numerical && env->getLogger())
        env->getLogger()->log(LOG_INFO, "LP relaxation flagged as infeasible or numerically unstable");

    if (result != nullptr)
    {
        result->status = numerical ? CODE_GENERIC_ERROR : CODE_SUCCESSFUL;
        result->terminationReason = static_cast<SolverTerminationReason>(r.reason);
        result->iterations = r.iterations;
        result->primalObj = r.primal_obj;
        result->dualObj = r.dual_obj;
        result->relativeGap = r.rel_gap;
    }
    if (!useWs) releaseIpmWorkspace(ws);
    return numerical ? CODE_GENERIC_ERROR : CODE_SUCCESSFUL;
}

SyphaStatus solver_sparse_mehrotra(SyphaNodeSparse &node)
{
    SolverExecutionConfig config;
    config.maxIterations = node.env->getMehrotraMaxIter();
    config.gapStagnation.enabled = false;
    config.bnbNodeOrdinal = 0;
    config.denseSelectionLogEveryNodes = 1;
    SolverExecutionResult result;
    SyphaStatus status = solver_sparse_mehrotra_run(node, config, &result);
    if (status != CODE_SUCCESSFUL) return status;
    return result.status;
}
