/* Stub: see gsl_matrix.h in this directory (src/sypha_solver_sparse.h:8). */
