"""TEST INFRASTRUCTURE (CPU oracle): restatement of sypha::Solver::Impl::buildStandardForm
(/root/reference/src/sypha_api.cpp:136-250) in plain Python loops - the general row model of the public API
(lb <= a.x <= ub per constraint, x >= 0, minimise or maximise) turned into the standard form the solver takes.
Only tests/ import this; the product's builder is sb200_build_standard_form (csrc/sb200_io.cu)."""
import math


def build_standard_form(n_vars, rows, lbs, ubs, objective, maximize=False):
    """rows: list of [(var, coef), ...] in the order the coefficients were set (Constraint::coeffs_, sypha_api.cpp:41-49);
    objective: {var: coef} (Objective::coeffs_).  Returns (nrows, ncols, csr_offs, csr_inds, csr_vals, obj, rhs)."""
    infos = []                                              # (constraint, is_ge, is_equality, rhs)  - sypha_api.cpp:153-186
    n_slacks = 0
    for ci, (lb, ub) in enumerate(zip(lbs, ubs)):
        has_lb, has_ub = math.isfinite(lb), math.isfinite(ub)
        if has_lb and has_ub and abs(lb - ub) <= 1e-15:
            infos.append((ci, True, True, lb))
        elif has_lb and has_ub:
            infos.append((ci, True, False, lb))
            infos.append((ci, False, False, ub))
            n_slacks += 2
        elif has_lb:
            infos.append((ci, True, False, lb))
            n_slacks += 1
        elif has_ub:
            infos.append((ci, False, False, ub))
            n_slacks += 1
        else:
            infos.append((ci, True, True, 0.0))
    nrows, ncols = len(infos), n_vars + n_slacks
    obj = [0.0] * ncols                                     # :189-197
    for var, coef in objective.items():
        if 0 <= var < n_vars:
            obj[var] = -coef if maximize else coef
    offs, inds, vals, rhs = [0], [], [], [0.0] * nrows
    slack = n_vars
    for ri, (ci, is_ge, is_eq, rhs_val) in enumerate(infos):  # :206-246
        if is_eq:
            for var, coef in rows[ci]:
                inds.append(var)
                vals.append(coef)
            rhs[ri] = rhs_val
        elif is_ge:
            for var, coef in rows[ci]:
                inds.append(var)
                vals.append(coef)
            inds.append(slack)
            vals.append(-1.0)
            slack += 1
            rhs[ri] = rhs_val
        else:
            for var, coef in rows[ci]:
                inds.append(var)
                vals.append(-coef)
            inds.append(slack)
            vals.append(-1.0)
            slack += 1
            rhs[ri] = -rhs_val
        offs.append(len(vals))
    return nrows, ncols, offs, inds, vals, obj, rhs
