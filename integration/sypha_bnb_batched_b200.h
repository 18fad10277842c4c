// sypha_bnb_batched_b200.h - the batched node loop offered next to the reference's
// solver_sparse_branch_and_bound (src/sypha_solver_bnb_driver.cpp:163): same inputs (a SyphaNodeSparse holding the host CSR of
// the instance), same outputs (node.objvalPrim / objvalDual / mipGap / iterations / hX in input-column space), K node LPs
// in flight on libsypha_b200 instead of one.
#pragma once
#include "common.h"

class SyphaNodeSparse;

struct SyphaBatchedBnbConfig
{
    int slots = 148;                     // node LPs per window (one thread block each, one launch per window; 148 = the SMs of a B200)
    int maxIterations = 0;               // per node LP; 0: env->getMehrotraMaxIter()
    int maxNodes = 0;                    // 0: env->getBnbMaxNodes()
    int maxDepth = 64;                   // branch decisions a workspace is sized for
    bool nodeLpToConvergence = false;    // false: the reference's gap-stagnation exit (bnb_driver.cpp:833-837)
    int windowsInFlight = 2;             // windows: sets of `slots` workspaces used alternately (sb200_window_begin / _finish), so the next
                                         // window is queued on the GPU while the host branches on the previous one; 1: one window at a time
    bool continuousBatching = false;     // true: sb200_solve_stream (a slot takes its next node at once) instead of windows of K
    bool referencePreprocessing = true;  // cost-driven and dominance reductions (bnb_driver.cpp:308-334)
};

struct SyphaBatchedBnbStats
{
    int processedNodes = 0, totalLpIterations = 0, prunedByBound = 0, failedLps = 0, integralNodes = 0, droppedTooDeep = 0;
    int inFlight = 0, openNodes = 0, baseRows = 0, baseColsOriginal = 0;
    double greedyIncumbent = 0.0, incumbent = 0.0, rootBound = 0.0, wallMs = 0.0, nodesPerSecond = 0.0, lpDeviceMs = 0.0;
};

SyphaStatus solver_sparse_branch_and_bound_batched(SyphaNodeSparse &node, const SyphaBatchedBnbConfig &cfg,
                                                   SyphaBatchedBnbStats *stats);
