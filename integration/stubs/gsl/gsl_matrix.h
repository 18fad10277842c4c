/* Stub: the B200 path has no GSL dependency.  The reference's sypha_solver_sparse.h still names this
 * header (src/sypha_solver_sparse.h:7); a maintainer adopting the shim deletes that include line, and
 * until then this empty file keeps the unchanged callers compiling. */
#define gsl_min(a, b) ((a) < (b) ? (a) : (b))
#define gsl_max(a, b) ((a) > (b) ? (a) : (b))
