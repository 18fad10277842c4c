/* oracle/refbuild/gsl/gsl_linalg.h - TEST INFRASTRUCTURE: everything is declared in gsl_matrix.h of this
 * directory (the reference includes both, src/sypha_solver_sparse.h:7-8). */
#include "gsl_matrix.h"
