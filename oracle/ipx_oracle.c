/* CPU restatement of IPX's per-iteration KKT solve -- TEST INFRASTRUCTURE ONLY.
 * See ipx_oracle.h for scope and parity status. Each function cites the
 * reference lines it follows; operation order is kept so that results are
 * bit-identical to the reference built with its own flags (-O2, no FMA).
 * Compile with -ffp-contract=off. */

#include "ipx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- building blocks: src/sparse_matrix.h:136-152, src/utils.cc:32-45 ---- */

static double dot_column(const oint* Ap, const oint* Ai, const double* Ax,
                         oint j, const double* rhs) {
    double d = 0.0;
    for (oint p = Ap[j]; p < Ap[j + 1]; p++) d += rhs[Ai[p]] * Ax[p];
    return d;
}

static void scatter_column(const oint* Ap, const oint* Ai, const double* Ax,
                           oint j, double alpha, double* lhs) {
    for (oint p = Ap[j]; p < Ap[j + 1]; p++) lhs[Ai[p]] += alpha * Ax[p];
}

static double dot(oint m, const double* x, const double* y) {
    double d = 0.0;
    for (oint i = 0; i < m; i++) d += x[i] * y[i];
    return d;
}

static double infnorm(oint m, const double* x) {
    double norm = 0.0;
    for (oint i = 0; i < m; i++) {
        double a = fabs(x[i]);
        if (a > norm) norm = a; /* std::max(norm, abs) */
    }
    return norm;
}

/* ---- NormalMatrix::_Apply, src/normal_matrix.cc:45-126 ---- */

void orc_normal_apply(oint m, oint n, const oint* Ap, const oint* Ai,
                      const double* Ax, const double* W, const double* rhs,
                      double* lhs, double* dotp) {
    if (W) {
        for (oint i = 0; i < m; i++) lhs[i] = rhs[i] * W[n + i]; /* :65-66 */
        for (oint j = 0; j < n; j++) {                           /* :67-75 */
            oint begin = Ap[j], end = Ap[j + 1];
            double d = 0.0;
            for (oint p = begin; p < end; p++) d += rhs[Ai[p]] * Ax[p];
            d *= W[j];
            for (oint p = begin; p < end; p++) lhs[Ai[p]] += d * Ax[p];
        }
    } else {
        for (oint i = 0; i < m; i++) lhs[i] = 0.0; /* :113 */
        for (oint j = 0; j < n; j++) {             /* :114-121 */
            oint begin = Ap[j], end = Ap[j + 1];
            double d = 0.0;
            for (oint p = begin; p < end; p++) d += rhs[Ai[p]] * Ax[p];
            for (oint p = begin; p < end; p++) lhs[Ai[p]] += d * Ax[p];
        }
    }
    if (dotp) *dotp = dot(m, rhs, lhs); /* :123-124 */
}

/* ---- DiagonalPrecond, src/diagonal_precond.cc:28-46, 150-157 ---- */

void orc_diag_build(oint m, oint n, const oint* Ap, const oint* Ai,
                    const double* Ax, const double* W, double* diag) {
    if (W) {
        for (oint i = 0; i < m; i++) diag[i] = W[n + i];
        for (oint j = 0; j < n; j++) {
            double w = W[j];
            for (oint p = Ap[j]; p < Ap[j + 1]; p++)
                diag[Ai[p]] += Ax[p] * w * Ax[p];
        }
    } else {
        for (oint i = 0; i < m; i++) diag[i] = 0.0;
        for (oint j = 0; j < n; j++)
            for (oint p = Ap[j]; p < Ap[j + 1]; p++)
                diag[Ai[p]] += Ax[p] * Ax[p];
    }
}

void orc_diag_apply(oint m, const double* diag, const double* rhs, double* lhs,
                    double* dotp) {
    double rldot = 0.0;
    for (oint i = 0; i < m; i++) {
        lhs[i] = rhs[i] / diag[i];
        rldot += lhs[i] * rhs[i];
    }
    if (dotp) *dotp = rldot;
}

/* ---- AddNormalProduct, src/sparse_matrix.cc:211-222 ---- */

void orc_add_normal_product(oint nrow, oint ncol, const oint* Ap,
                            const oint* Ai, const double* Ax, const double* D,
                            const double* rhs, double* lhs) {
    (void)nrow;
    for (oint j = 0; j < ncol; j++) {
        double temp = dot_column(Ap, Ai, Ax, j, rhs);
        if (D) temp *= D[j] * D[j];
        scatter_column(Ap, Ai, Ax, j, temp, lhs);
    }
}

/* ---- MultiplyAdd, src/sparse_matrix.cc:194-209 ---- */

void orc_multiply_add(oint nrow, oint ncol, const oint* Ap, const oint* Ai,
                      const double* Ax, const double* rhs, double alpha,
                      double* lhs, char trans) {
    (void)nrow;
    if (trans == 't' || trans == 'T') {
        for (oint j = 0; j < ncol; j++)                              /* :201-202 */
            lhs[j] += alpha * dot_column(Ap, Ai, Ax, j, rhs);
    } else {
        for (oint j = 0; j < ncol; j++)                              /* :206-207 */
            scatter_column(Ap, Ai, Ax, j, alpha * rhs[j], lhs);
    }
}

/* ---- TriangularSolve, src/sparse_matrix.cc:224-301 ---- */

oint orc_triangular_solve(oint ncol, const oint* Ap, const oint* Ai,
                          const double* Ax, double* x, char trans, char uplo,
                          int unitdiag) {
    oint nz = 0;
    const int upper = (uplo == 'u' || uplo == 'U');
    if (trans == 't' || trans == 'T') {
        if (upper) { /* :233-247 */
            for (oint i = 0; i < ncol; i++) {
                oint begin = Ap[i];
                oint end = Ap[i + 1] - (unitdiag ? 0 : 1);
                double d = 0.0;
                for (oint p = begin; p < end; p++) d += x[Ai[p]] * Ax[p];
                x[i] -= d;
                if (!unitdiag) x[i] /= Ax[end];
                if (x[i] != 0.0) nz++;
            }
        } else { /* :249-263 */
            for (oint i = ncol - 1; i >= 0; i--) {
                oint begin = Ap[i] + (unitdiag ? 0 : 1);
                oint end = Ap[i + 1];
                double d = 0.0;
                for (oint p = begin; p < end; p++) d += x[Ai[p]] * Ax[p];
                x[i] -= d;
                if (!unitdiag) x[i] /= Ax[begin - 1];
                if (x[i] != 0.0) nz++;
            }
        }
    } else {
        if (upper) { /* :267-281 */
            for (oint j = ncol - 1; j >= 0; j--) {
                oint begin = Ap[j];
                oint end = Ap[j + 1] - (unitdiag ? 0 : 1);
                if (!unitdiag) x[j] /= Ax[end];
                double temp = x[j];
                if (temp != 0.0) {
                    for (oint p = begin; p < end; p++) x[Ai[p]] -= Ax[p] * temp;
                    nz++;
                }
            }
        } else { /* :283-297 */
            for (oint j = 0; j < ncol; j++) {
                oint begin = Ap[j] + (unitdiag ? 0 : 1);
                oint end = Ap[j + 1];
                if (!unitdiag) x[j] /= Ax[begin - 1];
                double temp = x[j];
                if (temp != 0.0) {
                    for (oint p = begin; p < end; p++) x[Ai[p]] -= Ax[p] * temp;
                    nz++;
                }
            }
        }
    }
    return nz;
}

void orc_forward_solve(oint dim, const oint* Lp, const oint* Li,
                       const double* Lx, const oint* Up, const oint* Ui,
                       const double* Ux, double* x) { /* :303-306 */
    orc_triangular_solve(dim, Lp, Li, Lx, x, 'n', 'l', 1);
    orc_triangular_solve(dim, Up, Ui, Ux, x, 'n', 'u', 0);
}

void orc_backward_solve(oint dim, const oint* Lp, const oint* Li,
                        const double* Lx, const oint* Up, const oint* Ui,
                        const double* Ux, double* x) { /* :308-311 */
    orc_triangular_solve(dim, Up, Ui, Ux, x, 't', 'u', 0);
    orc_triangular_solve(dim, Lp, Li, Lx, x, 't', 'l', 1);
}

/* ---- SplittedNormalMatrix::_Apply, src/splitted_normal_matrix.cc:90-117 ---- */

void orc_split_apply(const orc_split* S, const double* rhs, double* lhs,
                     double* work, double* dotp) {
    const oint m = S->dim;
    memcpy(work, rhs, (size_t)m * sizeof(double));                 /* :96 */
    orc_backward_solve(m, S->Lp, S->Li, S->Lx, S->Up, S->Ui, S->Ux, work);
    for (oint i = 0; i < m; i++) lhs[i] = 0.0;                     /* :102 */
    orc_add_normal_product(m, S->ncolN, S->Np, S->Ni, S->Nx, NULL, work, lhs);
    orc_forward_solve(m, S->Lp, S->Li, S->Lx, S->Up, S->Ui, S->Ux, lhs);
    for (oint i = 0; i < m; i++) lhs[i] += rhs[i];                 /* :112 */
    for (oint k = 0; k < S->num_free; k++) lhs[S->free_positions[k]] = 0.0;
    if (dotp) *dotp = dot(m, rhs, lhs);
}

static void op_apply(const orc_operator* op, const double* rhs, double* lhs,
                     double* dotp) {
    if (op->kind == 0)
        orc_normal_apply(op->m, op->n, op->Ap, op->Ai, op->Ax, op->W, rhs, lhs,
                         dotp);
    else
        orc_split_apply(op->split, rhs, lhs, op->work, dotp);
}

static double scaled_resnorm(oint m, const double* resscale,
                             const double* residual) {
    if (!resscale) return infnorm(m, residual);
    double resnorm = 0.0;
    for (oint i = 0; i < m; i++) {
        double a = fabs(resscale[i] * residual[i]);
        if (a > resnorm) resnorm = a;
    }
    return resnorm;
}

/* ---- ConjugateResiduals::Solve (unpreconditioned),
 *      src/conjugate_residuals.cc:14-88 ---- */

oint orc_cr_solve(const orc_operator* op, oint m, const double* rhs, double tol,
                  const double* resscale, oint maxiter, double* lhs,
                  oint* iter_out, double* hist, oint hist_cap) {
    double* residual = calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    double* step = calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    double* Cresidual = calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    double* Cstep = calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    double cdot = 0.0;
    oint errflag = 0, iter = 0;
    if (maxiter < 0) maxiter = m + 100;

    if (infnorm(m, lhs) == 0.0) { /* :33-38 */
        memcpy(residual, rhs, (size_t)m * sizeof(double));
    } else {
        op_apply(op, lhs, residual, NULL);
        for (oint i = 0; i < m; i++) residual[i] = rhs[i] - residual[i];
    }
    op_apply(op, residual, Cresidual, &cdot);
    memcpy(step, residual, (size_t)m * sizeof(double));
    memcpy(Cstep, Cresidual, (size_t)m * sizeof(double));

    for (;;) {
        double resnorm = scaled_resnorm(m, resscale, residual); /* :44-49 */
        if (hist && iter < hist_cap) hist[iter] = resnorm;
        if (resnorm <= tol) break;
        if (iter == maxiter) { errflag = ORC_ERROR_cr_iter_limit; break; }
        if (cdot <= 0.0) { errflag = ORC_ERROR_cr_matrix_not_posdef; break; }
        const double denom = dot(m, Cstep, Cstep); /* :66 */
        const double alpha = cdot / denom;
        if (!isfinite(alpha)) { errflag = ORC_ERROR_cr_inf_or_nan; break; }
        for (oint i = 0; i < m; i++) lhs[i] += alpha * step[i];
        for (oint i = 0; i < m; i++) residual[i] -= alpha * Cstep[i];
        double cdotnew;
        op_apply(op, residual, Cresidual, &cdotnew); /* :75 */
        const double beta = cdotnew / cdot;
        for (oint i = 0; i < m; i++) step[i] = residual[i] + beta * step[i];
        for (oint i = 0; i < m; i++) Cstep[i] = Cresidual[i] + beta * Cstep[i];
        cdot = cdotnew;
        iter++;
    }
    free(residual); free(step); free(Cresidual); free(Cstep);
    if (iter_out) *iter_out = iter;
    return errflag;
}

/* ---- ConjugateResiduals::Solve (preconditioned),
 *      src/conjugate_residuals.cc:90-213 ---- */

oint orc_pcr_solve(const orc_operator* op, oint m, const double* diag,
                   const double* rhs, double tol, const double* resscale,
                   oint maxiter, double* lhs, oint* iter_out, double* hist,
                   oint hist_cap) {
    size_t cnt = (size_t)(m > 0 ? m : 1);
    double* residual = calloc(cnt, sizeof(double));
    double* sresidual = calloc(cnt, sizeof(double));
    double* step = calloc(cnt, sizeof(double));
    double* Csresidual = calloc(cnt, sizeof(double));
    double* Cstep = calloc(cnt, sizeof(double));
    double cdot = 0.0;
    double resnorm_precond_system = 0.0;
    oint errflag = 0, iter = 0;
    if (maxiter < 0) maxiter = m + 100;

    if (infnorm(m, lhs) == 0.0) { /* :118-123 */
        memcpy(residual, rhs, (size_t)m * sizeof(double));
    } else {
        op_apply(op, lhs, residual, NULL);
        for (oint i = 0; i < m; i++) residual[i] = rhs[i] - residual[i];
    }
    orc_diag_apply(m, diag, residual, sresidual, &resnorm_precond_system);
    op_apply(op, sresidual, Csresidual, &cdot);
    memcpy(step, sresidual, (size_t)m * sizeof(double));
    memcpy(Cstep, Csresidual, (size_t)m * sizeof(double));

    for (;;) {
        double resnorm = scaled_resnorm(m, resscale, residual); /* :131-136 */
        if (hist && iter < hist_cap) hist[iter] = resnorm;
        if (resnorm <= tol) break;
        if (iter == maxiter) { errflag = ORC_ERROR_cr_iter_limit; break; }
        if (cdot <= 0.0) { errflag = ORC_ERROR_cr_matrix_not_posdef; break; }
        double cdotnew;
        {
            double* precond_Cstep = Csresidual; /* :160-162 aliasing */
            double pdot;
            orc_diag_apply(m, diag, Cstep, precond_Cstep, &pdot);
            if (pdot <= 0.0) { errflag = ORC_ERROR_cr_precond_not_posdef; break; }
            const double alpha = cdot / pdot;
            if (!isfinite(alpha)) { errflag = ORC_ERROR_cr_inf_or_nan; break; }
            for (oint i = 0; i < m; i++) lhs[i] += alpha * step[i];
            for (oint i = 0; i < m; i++) residual[i] -= alpha * Cstep[i];
            for (oint i = 0; i < m; i++) sresidual[i] -= alpha * precond_Cstep[i];
            op_apply(op, sresidual, Csresidual, &cdotnew); /* :176 */
        }
        const double beta = cdotnew / cdot;
        for (oint i = 0; i < m; i++) step[i] = sresidual[i] + beta * step[i];
        for (oint i = 0; i < m; i++) Cstep[i] = Csresidual[i] + beta * Cstep[i];
        cdot = cdotnew;
        iter++;
        if (iter % 5 == 0) { /* :187-207 */
            double rsdot;
            orc_diag_apply(m, diag, residual, sresidual, &rsdot);
            if (rsdot >= resnorm_precond_system) {
                errflag = ORC_ERROR_cr_no_progress;
                break;
            }
            resnorm_precond_system = rsdot;
        }
    }
    free(residual); free(sresidual); free(step); free(Csresidual); free(Cstep);
    if (iter_out) *iter_out = iter;
    return errflag;
}

/* ---- KKTSolverDiag, src/kkt_solver_diag.cc ---- */

void orc_kktdiag_weights(oint m, oint n, int have_iterate, const double* xl,
                         const double* xu, const double* zl, const double* zu,
                         double mu, double* W, double* resscale) {
    if (have_iterate) { /* :24-49 */
        double regval = mu;
        for (oint j = 0; j < n + m; j++) {
            double g = zl[j] / xl[j] + zu[j] / xu[j];
            if (g != 0.0 && g < regval) regval = g;
            W[j] = 1.0 / g;
        }
        for (oint j = 0; j < n + m; j++)
            if (isinf(W[j])) W[j] = 1.0 / regval;
    } else { /* :50-52 */
        for (oint j = 0; j < n + m; j++) W[j] = 1.0;
    }
    for (oint i = 0; i < m; i++) resscale[i] = 1.0 / sqrt(W[n + i]); /* :55-56 */
}

void orc_kktdiag_rhs(oint m, oint n, const oint* Ap, const oint* Ai,
                     const double* Ax, const double* W, const double* a,
                     const double* b, double* rhs) { /* :90-92 */
    for (oint i = 0; i < m; i++) rhs[i] = -b[i];
    for (oint j = 0; j < n + m; j++)
        scatter_column(Ap, Ai, Ax, j, W[j] * a[j], rhs);
}

void orc_kktdiag_recover(oint m, oint n, const oint* Ap, const oint* Ai,
                         const double* Ax, const double* W, const double* a,
                         const double* b, const double* y, double* x) {
    for (oint i = 0; i < m; i++) x[n + i] = b[i]; /* :108-109 */
    for (oint j = 0; j < n; j++) {                /* :110-117 */
        double aty = dot_column(Ap, Ai, Ax, j, y);
        x[j] = W[j] * (a[j] - aty);
        for (oint p = Ap[j]; p < Ap[j + 1]; p++) x[n + Ai[p]] -= x[j] * Ax[p];
    }
}

oint orc_kktdiag_solve(oint m, oint n, const oint* Ap, const oint* Ai,
                       const double* Ax, const double* W, const double* diag,
                       const double* resscale, const double* a, const double* b,
                       double tol, oint maxiter, double* x, double* y,
                       oint* iter) {
    double* rhs = malloc((size_t)(m > 0 ? m : 1) * sizeof(double));
    orc_kktdiag_rhs(m, n, Ap, Ai, Ax, W, a, b, rhs);
    for (oint i = 0; i < m; i++) y[i] = 0.0; /* :95 */
    orc_operator op;
    memset(&op, 0, sizeof op);
    op.kind = 0; op.m = m; op.n = n; op.Ap = Ap; op.Ai = Ai; op.Ax = Ax; op.W = W;
    oint err = orc_pcr_solve(&op, m, diag, rhs, tol, resscale, maxiter, y, iter,
                             NULL, 0);
    orc_kktdiag_recover(m, n, Ap, Ai, Ax, W, a, b, y, x);
    free(rhs);
    return err;
}
