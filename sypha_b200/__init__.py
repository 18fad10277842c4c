"""sypha_b200 - B200-native (sm_100a) Mehrotra IPM hot path behind sypha's solver interface.

Only what the path needs: ``csrc/`` (CUDA kernels + the C ABI of include/sypha_b200.h) and
``solver.py`` (host-side mirror of the reference's operator interface over that ABI).
"""
from .solver import (  # noqa: F401
    CODE_GENERIC_ERROR, CODE_SUCCESSFUL, IpmWorkspace, Sb200Error, SolverExecutionConfig,
    SolverExecutionResult, SolverGapStagnationConfig, SyphaEnvironment, SyphaNodeSparse,
    SOLVER_TERM_CONVERGED, SOLVER_TERM_GAP_STALLED, SOLVER_TERM_INFEASIBLE_OR_NUMERICAL,
    SOLVER_TERM_MAX_ITER, SOLVER_TERM_TIME_LIMIT, initializeIpmWorkspace, releaseIpmWorkspace,
    solve_batch, solver_sparse_mehrotra, solver_sparse_mehrotra_run)
