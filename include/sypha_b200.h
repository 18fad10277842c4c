/*
 * sypha_b200.h - C ABI of the B200-native (sm_100a) Mehrotra IPM hot path.
 *
 * This is the drop-in boundary for the path BASELINE.json's north_star names: the
 * per-iteration linear algebra of sypha's Mehrotra predictor-corrector LP solve.  Every entry
 * point is `extern "C"`, takes plain pointers and sizes, returns an int status, never throws
 * and never calls exit().  Reference interfaces replaced (paths under /root/reference):
 *
 *   sb200_solve ................ solver_sparse_mehrotra_run          src/sypha_solver_sparse.h:51
 *                                (body src/sypha_solver.cpp:42-886)
 *   sb200_ws_create/destroy .... initializeIpmWorkspace / releaseIpmWorkspace
 *                                src/sypha_solver.h:107-108, src/sypha_solver_workspace.cpp:5-89
 *   sb200_load_model ........... SyphaNodeSparse::copyModelOnDevice   src/sypha_node_sparse.cpp:156-198
 *                                + host KKT assembly src/sypha_solver.cpp:96-207 (disappears)
 *   sb200_solve_batch .......... the B&B node body                    src/sypha_solver_bnb_driver.cpp:789-859
 *                                + build_branch_model                 src/sypha_solver_bnb.cpp:418-490
 *   sb200_node_heuristics ...... MostFractionalSelector, NearestIntegerFixingHeuristic and the cover repair
 *                                of DualGuidedCoverRepairHeuristic (plain versions), run per node by
 *                                src/sypha_solver_bnb_driver.cpp:861-1005 on a host copy of the LP point
 *                                src/sypha_solver_heuristics.cpp:10-30,53-110,112-292
 *   sb200_k_* .................. the L0 free functions (device pointers + stream):
 *     sb200_k_elem_min_mult .... elem_min_mult_dev                    src/sypha_solver_utils.h:7
 *     sb200_k_corrector_rhs .... corrector_rhs_dev                    src/sypha_solver_utils.h:12
 *     sb200_k_alpha_max ........ alpha_max_dev                        src/sypha_solver_utils.h:17-24
 *     sb200_k_spmv_csr/_csc .... cusparseSpMV (NON_TRANSPOSE / TRANSPOSE), call sites SURVEY.md 2.2
 *     sb200_k_jacobi_diag ...... krylovComputeJacobiDiag              src/sypha_solver_krylov.h:49
 *     sb200_k_potrf/_potrs ..... cusolverDnDgetrf / Dgetrs            src/sypha_solver_dense_linear.cpp:179,195
 *     sb200_k_syrk ............. (new) FP64 tensor-core A*diag(d)*A' on a dense A
 *
 * All arithmetic is FP64, all indices int32 (as in the reference).  Device pointers are raw CUDA
 * device addresses in the current context; `stream` arguments are a `cudaStream_t` passed as void*.
 */
#ifndef SYPHA_B200_H
#define SYPHA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_VERSION 100

typedef struct sb200_ws sb200_ws;

/* status codes (return values) */
enum {
    SB200_OK = 0,
    SB200_ERR_INVALID = 1,     /* bad argument / model not loaded / capacity exceeded */
    SB200_ERR_CUDA = 2,        /* a CUDA call failed; see sb200_last_error */
    SB200_ERR_NOMEM = 3,
    SB200_ERR_NUMERICAL = 4,   /* LP flagged infeasible-or-numerical (maps to CODE_GENERIC_ERROR) */
    SB200_ERR_UNSUPPORTED = 5
};

/* termination reasons: same values as SolverTerminationReason, src/sypha_solver_sparse.h:13-20 */
enum {
    SB200_TERM_CONVERGED = 0,
    SB200_TERM_MAX_ITER = 1,
    SB200_TERM_GAP_STALLED = 2,
    SB200_TERM_INFEASIBLE_OR_NUMERICAL = 3,
    SB200_TERM_TIME_LIMIT = 4
};

/* linear-solve strategy (extends the reference's auto|dense|sparse_qr|krylov,
 * src/sypha_environment_defaults.h:26-30) */
enum {
    SB200_STRATEGY_AUTO = 0,
    SB200_STRATEGY_CHOLESKY = 1,  /* sparse symbolic/numeric assembly of M = A D A' + dense Cholesky */
    SB200_STRATEGY_SYRK = 2,      /* FP64 tensor-core SYRK on a dense copy of A + dense Cholesky */
    SB200_STRATEGY_PCG = 3        /* matrix-free Jacobi-PCG on the normal equations */
};

/* capacities of a persistent workspace (grow-only; 0 = size on first load) */
typedef struct sb200_caps {
    int m_max;
    int n_max;
    long long nnz_max;
} sb200_caps;

/* parameters of one LP solve; defaults = src/sypha_environment_defaults.h:14-24 */
typedef struct sb200_params {
    int max_iter;               /* kMehrotraMaxIter = 25 */
    double eta;                 /* kMehrotraEta = 0.95 */
    double mu_tol;              /* kMehrotraMuTol = 1e-4 */
    int gap_enabled;            /* SolverGapStagnationConfig, src/sypha_solver_sparse.h:22-27 */
    int gap_window;
    double gap_min_improv_pct;
    int strategy;               /* SB200_STRATEGY_* */
    int cg_max_iter;            /* kKrylovMaxCgIter = 500 */
    double cg_tol_initial;      /* 1e-2 */
    double cg_tol_final;        /* 1e-8 */
    double cg_tol_decay;        /* 0.5 */
    const volatile int *stop_flag;  /* host flag polled between iterations (logger watchdog), may be NULL */
    int poll_every;             /* host reads the device scalar block every k iterations (>=1) */
    int use_graph;              /* 1 = replay one captured CUDA graph per iteration */
    /* live stop test, asked every time the host looks at the scalar block (every poll_every iterations): the
     * reference polls node.env->getLogger()->isStopRequested() once per iteration (src/sypha_solver.cpp:498-502),
     * an std::atomic<bool> behind a getter that no plain int pointer can mirror.  May be NULL. */
    int (*stop_cb)(void *user);
    void *stop_user;
} sb200_params;

typedef struct sb200_result {
    int status;                 /* SB200_OK or SB200_ERR_NUMERICAL */
    int reason;                 /* SB200_TERM_* */
    int iterations;
    double primal_obj;          /* x[0:n_orig].c[0:n_orig]   (src/sypha_solver.cpp:781) */
    double dual_obj;            /* y.b                        (:784) */
    double rel_gap;             /* |p-d|/max(1,|p|)           (:786-787) */
    double mu;
    double ms_start;            /* starting point     (node.timeStartSol*) */
    double ms_setup;            /* initial residuals  (node.timePreSol*)   */
    double ms_loop;             /* main loop          (node.timeSolver*)   */
    int strategy_used;
    long long cg_iterations;    /* total CG iterations (PCG strategy) */
    long long kernels_launched; /* kernels of this library launched by the call */
    double *x_host;             /* out, length n, may be NULL */
    double *y_host;             /* out, length m, may be NULL */
    double *s_host;             /* out, length n, may be NULL */
    double *x0_host;            /* out: starting point (node.hX/hY/hS), may be NULL */
    double *y0_host;
    double *s0_host;
    double *xys_device;         /* out, DEVICE memory, may be NULL: the final x[n] | y[m] | s[n] packed - what a child node
                                   of the B&B takes as its warm start (sb200_node_delta.warm_start) */
} sb200_result;

/* one B&B node = base model + appended rows (build_branch_model, src/sypha_solver_bnb.cpp:453-468):
 * row r has coefficient `coef[r]` at column `var[r]` and -1 at a fresh slack column, rhs `rhs[r]`. */
typedef struct sb200_node_delta {
    int n_extra_rows;
    const int *var;
    const double *coef;
    const double *rhs;
    /* Warm start of a child LP from its parent's iterate (the reference starts every node cold,
     * src/sypha_solver.cpp:77).  warm_start: DEVICE pointer to the parent's final x[warm_n] | y[warm_m] | s[warm_n]
     * (sb200_result.xys_device of the parent's solve; the parent's rows and columns are the first warm_m / warm_n of the
     * child) or NULL for the Mehrotra starting point.  The child starts from x = max(x_parent, warm_floor),
     * s = max(s_parent, warm_floor), y = y_parent, new rows and columns at warm_floor / 0 - the parent's optimal
     * face pulled back into the interior - and skips the starting-point factorisation.  Honoured by the throughput form
     * (SB200_FORM_THROUGHPUT); the latency form starts cold. */
    const double *warm_start;
    int warm_n, warm_m;
    double warm_floor;          /* <= 0: 0.1 */
    double *export_xys;         /* DEVICE, may be NULL: receives this node's final x | y | s (sb200_solve_batch /
                                   sb200_solve_stream set sb200_result.xys_device from it) - the warm start of its children */
} sb200_node_delta;

/* ---- lifetime ------------------------------------------------------------------------------ */
int sb200_version(void);
int sb200_device_count(void);
int sb200_ws_create(int device, const sb200_caps *caps, sb200_ws **out);
int sb200_ws_destroy(sb200_ws *ws);
const char *sb200_last_error(const sb200_ws *ws);
void sb200_default_params(sb200_params *p);

/* ---- model ---------------------------------------------------------------------------------- */
/* A is m x n CSR; the first n_orig columns carry the objective that is reported.  Builds the CSC
 * copy and, for the direct strategies, the symbolic structure of M = A D A' once.
 * `strategy_hint` = SB200_STRATEGY_*; AUTO decides from the instance. */
int sb200_load_model(sb200_ws *ws, int m, int n, int n_orig, long long nnz,
                     const int *csr_offs, const int *csr_inds, const double *csr_vals,
                     const double *c, const double *b, int ptrs_on_device, int strategy_hint);

/* Host side of the model: an OR-Library set-covering text file read straight into the standard form
 * A = [A0 | -I] (surplus entry last in its row), b = 1, c = [c0; 0] - what
 * model_reader_read_scp_file_sparse_csr builds (src/model_reader.cpp:90-174), in one pass over the file.  The
 * arrays are owned by the library until sb200_free_scp; they are what sb200_load_model takes. */
typedef struct sb200_scp_model {
    int m, n, n_orig;
    long long nnz;
    int *csr_offs;       /* [m + 1] */
    int *csr_inds;       /* [nnz] */
    double *csr_vals;    /* [nnz] */
    double *c;           /* [n] */
    double *b;           /* [m] */
} sb200_scp_model;
int sb200_read_scp(const char *path, sb200_scp_model *out);

/* A general row model (lb <= a.x <= ub per row, x >= 0, minimise or maximise c.x) -> the standard form the solver takes
 * (every row an equality, surplus columns appended): what sypha::Solver::Impl::buildStandardForm does
 * (src/sypha_api.cpp:136-250), row for row and column for column - an equality row (lb == ub) keeps its coefficients; a
 * ">= lb" row gains a surplus column with -1; a "<= ub" row is negated (coefficients and right-hand side) and gains a
 * surplus column with -1; a range row becomes those two rows in that order; a row without bounds is kept as "= 0".
 * Surplus columns are numbered in row order behind the n_vars structural ones; their cost is 0; a maximised objective is
 * negated.  O(nnz): a caller that holds its rows as arrays needs neither Variable / Constraint objects nor their
 * O(row length) SetCoefficient.  Host code; the outputs are what sb200_load_model takes. */
typedef struct sb200_row_model {
    int n_vars, n_rows;
    const int *row_offs;     /* [n_rows + 1] */
    const int *row_inds;     /* variable of each coefficient, in the order the coefficients were set */
    const double *row_vals;
    const double *row_lb;    /* [n_rows], -INFINITY: none */
    const double *row_ub;    /* [n_rows], +INFINITY: none */
    const double *obj;       /* [n_vars] */
    int maximize;
} sb200_row_model;
/* sizes of the standard form: rows, columns (structural + surplus), stored entries */
int sb200_standard_form_size(const sb200_row_model *in, int *nrows, int *ncols, long long *nnz);
/* fills caller-owned arrays of those sizes: csr_offs[nrows + 1], csr_inds[nnz], csr_vals[nnz], obj[ncols], rhs[nrows] */
int sb200_build_standard_form(const sb200_row_model *in, int *csr_offs, int *csr_inds, double *csr_vals, double *obj, double *rhs);
void sb200_free_scp(sb200_scp_model *mdl);

/* ---- solve ---------------------------------------------------------------------------------- */
int sb200_solve(sb200_ws *ws, const sb200_params *params, sb200_result *result);
/* Turn the resident BASE model into the model of one B&B node (base + appended branch rows,
 * build_branch_model, src/sypha_solver_bnb.cpp:418-490) ON THE DEVICE: no host CSR copy, no re-upload, no
 * new symbolic structure (replaces the per-node build + SyphaNodeSparse::copyModelOnDevice of
 * src/sypha_solver_bnb_driver.cpp:807-826).  delta == NULL or n_extra_rows == 0 restores the base model.
 * The workspace must have been created with sb200_caps covering the deepest node (m, n, nnz + 2 per row);
 * direct (sparse-assembly + Cholesky) strategy only, otherwise SB200_ERR_UNSUPPORTED.  The next
 * sb200_solve / x_host, y_host, s_host have the node's dimensions (m + k, n + k). */
int sb200_set_node_delta(sb200_ws *ws, const sb200_node_delta *delta);
/* k independent LPs, one workspace (own stream) each, solved concurrently.  deltas == NULL: the
 * workspaces' resident models as they are; otherwise deltas[i] is applied to wss[i] first.
 * With every workspace on one device and in the throughput form (sb200_set_solver_form / sb200_set_concurrency_hint)
 * the batch is a WINDOW: all deltas are staged in one host->device copy and applied by one kernel, the k LPs are ONE
 * launch of k thread blocks on wss[0]'s stream (any k; 148 = one block per SM of a B200; more blocks queue in the
 * hardware scheduler and start as SMs free up), results[i].x_host / y_host / s_host are filled behind it.
 * sb200_last_window(wss[0], ...) then reports the device time of that launch. */
int sb200_solve_batch(sb200_ws **ws, int k, const sb200_node_delta *deltas,
                      const sb200_params *params, sb200_result *results);

/* The combinatorial step that follows a node's LP in the B&B loop, on the device (one single-CTA kernel per
 * node on the node's stream, reading the LP point where the solve left it): most-fractional branching variable
 * (src/sypha_solver_heuristics.cpp:10-30), nearest-integer rounding with greedy cover repair and removal of
 * redundant columns for an incumbent (:53-110, :112-292, plain versions; columns the node's decisions fix to 0
 * are never chosen), and c.rint(x) for an integral LP point.  Only the first n_orig columns and the base rows
 * take part. */
typedef struct sb200_heur_result {
    int feasible;               /* 0: no cover exists under the node's zero-fixings */
    int n_chosen;               /* columns in the cover */
    int branch_var;             /* most fractional original column, first index on ties (-1: none) */
    int repair_steps;           /* greedy picks that were needed */
    double cover_obj;           /* cost of the cover (DBL_MAX when infeasible) */
    double branch_frac;         /* |x - rint(x)| at branch_var */
    double rounded_obj;         /* c . rint(x) over the original columns */
    int nif_feasible;           /* SB200_HEUR_REFERENCE: NearestIntegerFixingHeuristic (rounding + decisions, no repair) covers */
    int reserved;
    double nif_obj;             /* its cost (DBL_MAX when it does not cover) */
} sb200_heur_result;
/* Which rules sb200_node_heuristics runs.
 *   SB200_HEUR_REFERENCE (default): the reference's own - fractional candidates by `tol` and the selector
 *     `branch_rule` (collect_fractional_candidates src/sypha_solver_bnb.cpp:368-382; MostFractionalSelector /
 *     HighestCostFractionalSelector src/sypha_solver_heuristics.cpp:10-51), NearestIntegerFixingHeuristic (:53-110)
 *     -> nif_feasible / nif_obj / sb200_get_rounded, DualGuidedCoverRepairHeuristic (:112-292, dual guidance from the
 *     resident y) -> feasible / cover_obj / n_chosen / repair_steps / sb200_get_cover.  Set-covering models only (unit
 *     coefficients in the base rows, rhs 1): anything else comes back as repair_steps = -1 with no cover.
 *   SB200_HEUR_PLAIN: round at 0.5, greedy repair by cost per newly covered row, redundant columns dropped. */
enum { SB200_HEUR_PLAIN = 0, SB200_HEUR_REFERENCE = 1 };
enum { SB200_BRANCH_MOST_FRACTIONAL = 0, SB200_BRANCH_HIGHEST_COST_FRACTIONAL = 1 };
int sb200_set_heuristic_rules(sb200_ws *ws, int rules, int branch_rule, double integrality_tol);
/* runs the kernel for each of the k workspaces (their last solve must have finished) concurrently, then waits */
int sb200_node_heuristics(sb200_ws **ws, int k, sb200_heur_result *out);
/* the cover of the last sb200_node_heuristics on this workspace: n_orig bytes of 0/1 */
int sb200_get_cover(sb200_ws *ws, unsigned char *x_host);
/* the NearestIntegerFixing rounding of the last sb200_node_heuristics (SB200_HEUR_REFERENCE): n_orig bytes of 0/1 */
int sb200_get_rounded(sb200_ws *ws, unsigned char *x_host);

/* Throughput hint for workspaces that solve LPs CONCURRENTLY (B&B slots): the data-flow factorisation then
 * launches about (2 x SMs) / concurrent_lps CTAs instead of one per task - a CTA that waits for a dependency
 * holds its SM slot, which is free latency hiding for one LP and lost throughput for sixteen.  0 or 1 restores
 * the single-LP geometry.  Cached iteration graphs are dropped. */
int sb200_set_concurrency_hint(sb200_ws *ws, int concurrent_lps);

/* Shape of the solve.  SB200_FORM_LATENCY (default): one LP spread over the whole GPU, ~12 kernels per iteration
 * replayed from a CUDA graph - the fastest way to ONE optimum.  SB200_FORM_THROUGHPUT: the whole LP (starting point,
 * every iteration, termination test) is ONE launch of ONE thread block (sb200_cta.cu) - several times slower for one
 * LP, but 148 of them run side by side with no launch, host round trip or inter-block wait per phase: the form for
 * B&B node LPs and batched relaxations (sb200_set_concurrency_hint with more than one LP selects it).  Needs the
 * sparse-assembly + Cholesky strategy on a unit-coefficient model with at most 2048 rows (otherwise the latency form
 * runs); the starting point is not copied out (x0_host / y0_host / s0_host must be NULL) and the stop flag is not
 * polled inside an LP. */
enum { SB200_FORM_LATENCY = 0, SB200_FORM_THROUGHPUT = 1 };
int sb200_set_solver_form(sb200_ws *ws, int form);
/* sb200_solve_batch over workspaces in the throughput form is ONE launch (a thread block per LP) on the first workspace's
 * stream.  Device time of the last such window on `ws` (= the first workspace of that batch), measured with CUDA events on
 * the launching stream around the kernel, and the number of LPs it held. */
int sb200_last_window(sb200_ws *ws, double *ms, int *lps);
/* Optional, once per base model (after sb200_load_model, before the first node): allocate what the node path otherwise
 * allocates at a workspace's first node - the copy of the base CSC, the delta arrays for max_extra_rows decisions, the
 * buffers of sb200_node_heuristics - so that no allocation (an implicit device synchronisation) falls into the search,
 * where it would wait for a window in flight on another set of workspaces. */
int sb200_prepare_nodes(sb200_ws *ws, int max_extra_rows);
/* The same window in two halves, so that a caller can keep the GPU fed: begin applies the deltas, launches the window on
 * wss[0]'s stream and - with_node_rules != 0 - the window's node-rules kernel (sb200_node_heuristics) right behind it,
 * and returns without waiting; finish waits for that stream and fills results[] (the SAME array that was passed to begin:
 * its x_host / y_host / s_host / xys_device pointers are used by the copies queued in begin) and rules_out[] (may be NULL).
 * Two sets of workspaces used alternately - begin(A), begin(B), finish(A), ..., begin(A'), finish(B), ... - leave no gap
 * between windows: B's thread blocks start on the SMs A's finished blocks free.  SB200_ERR_UNSUPPORTED: the batch is not a
 * one-launch window (see sb200_solve_batch); any deltas have then been applied and sb200_solve_batch(wss, k, NULL, ...)
 * solves it. */
int sb200_window_begin(sb200_ws **wss, int k, const sb200_node_delta *deltas, const sb200_params *params,
                       sb200_result *results, int with_node_rules);
int sb200_window_finish(sb200_ws **wss, int k, sb200_result *results, sb200_heur_result *rules_out);

/* Continuous batching of B&B node LPs over k workspaces that hold the same base model: whenever a slot is
 * free `next(user, slot, &delta)` is asked for a node (return 1 with the decision list filled in - the arrays
 * are copied before `next` returns control a second time - or 0 if there is none right now); the node's LP is
 * solved (as sb200_set_node_delta + sb200_solve with x_host = y_host = s_host = NULL), sb200_node_heuristics
 * runs behind it, and `done(user, slot, result, heur)` hands both records back - it may add nodes to the
 * caller's frontier.  Returns when every slot is idle and `next` has nothing.  Callbacks run on the calling
 * thread.  Replaces the one-node-at-a-time loop of src/sypha_solver_bnb_driver.cpp:698-1046 without a window
 * barrier: a slot never waits for another slot's LP. */
typedef int (*sb200_next_node_fn)(void *user, int slot, sb200_node_delta *delta);
typedef void (*sb200_node_done_fn)(void *user, int slot, const sb200_result *result, const sb200_heur_result *heur);
int sb200_solve_stream(sb200_ws **ws, int k, const sb200_params *params, sb200_next_node_fn next,
                       sb200_node_done_fn done, void *user);

/* per-iteration trace of the last solve: rows of SB200_TRACE_COLS doubles
 * (mu_in, mu, mu_aff, sigma, alpha_p, alpha_d, primal, dual); returns rows written */
#define SB200_TRACE_COLS 8
int sb200_get_trace(sb200_ws *ws, double *out, int max_rows);
/* device addresses of the resident iterates (length n, m, n) */
int sb200_get_device_iterates(sb200_ws *ws, void **x, void **y, void **s);
/* host copies of the resident iterates (any pointer may be NULL) */
int sb200_get_iterates(sb200_ws *ws, double *x_host, double *y_host, double *s_host);
/* introspection used by bench.py for the roofline arithmetic */
int sb200_model_info(sb200_ws *ws, long long *info, int n_info);
void *sb200_stream(sb200_ws *ws);
/* time one phase of the loop on the resident model with CUDA events on the workspace stream:
 * phase 0 = normal-matrix assembly, 1 = Cholesky factorisation, 2 = one solve (forward+backward),
 * 3 = CSR SpMV (rhs), 4 = CSC SpMV + recovery + ratio test, 5 = fused step-length / mu_aff / sigma / corrector kernel(s),
 * 6 = one whole CG iteration, 7 = its A'p product, 8 = its A q product (PCG strategy only).
 * Writes the mean milliseconds per launch group over `reps` runs (after one warm-up). */
int sb200_time_phase(sb200_ws *ws, int phase, int reps, double *ms_out);

/* y = A x (transpose = 0, x length n, y length m) or y = A' x (transpose = 1) on the RESIDENT model, with
 * whatever representation the workspace built for it (value-carrying CSR/CSC, or the +/-1 pattern staged
 * through shared memory for PCG-strategy models).  Device pointers; runs on the workspace stream and
 * synchronises.  Replaces cusparseSpMV on node.matDescr (src/sypha_solver.cpp:419,450). */
int sb200_ws_spmv(sb200_ws *ws, int transpose, const double *d_x, double *d_y);

/* ---- L0 kernels on caller-owned device buffers ------------------------------------------------ */
int sb200_k_elem_min_mult(const double *d_x, const double *d_s, double *d_out, int n, void *stream);
int sb200_k_corrector_rhs(const double *d_dx, const double *d_ds, double sigma, double mu,
                          double *d_out, int n, void *stream);
/* d_result[0] = min_{dx<0} -x/dx, d_result[1] = min_{ds<0} -s/ds (DBL_MAX when empty); stays on the
 * device; if h_result != NULL it is also copied out (one sync), any n (the reference bails at
 * n > 262144, src/sypha_solver_utils.cu:151-154) */
int sb200_k_alpha_max(const double *d_x, const double *d_dx, const double *d_s, const double *d_ds,
                      int n, double *d_result, double *h_result, void *stream);
/* y = alpha*A*x + beta*y, A m x n CSR */
int sb200_k_spmv_csr(int m, const int *d_offs, const int *d_inds, const double *d_vals,
                     const double *d_x, double *d_y, double alpha, double beta, void *stream);
/* y = alpha*A'*x + beta*y given the CSC arrays of A (n columns) */
int sb200_k_spmv_csc(int n, const int *d_colptr, const int *d_rows, const double *d_vals,
                     const double *d_x, double *d_y, double alpha, double beta, void *stream);
/* diag[i] = sum_k a_ik^2 d[col_k] */
int sb200_k_jacobi_diag(int m, const int *d_offs, const int *d_inds, const double *d_vals,
                        const double *d_d, double *d_diag, void *stream);
/* in-place lower Cholesky of the row-major n x n matrix with leading dimension ld (ld % 64 == 0,
 * rows/cols n..ld-1 must hold an identity pad); *d_info = 0 or 1-based index of the first
 * non-positive pivot */
int sb200_k_potrf(int n, double *d_a, int ld, int *d_info, void *stream);
/* solve L L' x = b in place (b length ld, zero padded) */
int sb200_k_potrs(int n, const double *d_l, int ld, double *d_b, void *stream);
/* C (row-major, ld) lower triangle = A diag(d) A' for a dense row-major m x k matrix A (lda);
 * FP64 tensor-core (DMMA) kernel */
int sb200_k_syrk(int m, int k, const double *d_a, int lda, const double *d_d, double *d_c, int ld,
                 void *stream);
/* M (row-major, ld) lower triangle = A diag(d) A' through the workspace's symbolic structure */
int sb200_assemble_normal(sb200_ws *ws, const double *d_d, double *d_m, int ld);

#ifdef __cplusplus
}
#endif
#endif /* SYPHA_B200_H */
