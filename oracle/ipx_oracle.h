/* CPU restatement of IPX's per-iteration KKT solve -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (ipx_b200/csrc, ipx_b200/host) never does.
 *
 * Parity status: PINNED against the reference itself. Every function below is
 * checked bit-for-bit (same operation order, no FMA contraction) against the
 * reference's own classes compiled from /root/reference (oracle/_ref, recipe in
 * oracle/Makefile) by tests/test_oracle_vs_ref.py, and against the golden
 * vectors under tests/golden/ that were generated from that build
 * (tests/golden/make_golden.py). The reference's own test-suite holds no
 * numeric known-answer for this path.
 *
 * All matrices are CSC with int64 indices, as in the reference
 * (src/sparse_matrix.h, typedef int64_t ipxint in include/ipx_config.h:5).
 */
#ifndef IPX_ORACLE_H_
#define IPX_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t oint;

/* Error flags, reference include/ipx_status.h:31-35,47. */
#define ORC_ERROR_cr_iter_limit 201
#define ORC_ERROR_cr_matrix_not_posdef 202
#define ORC_ERROR_cr_precond_not_posdef 203
#define ORC_ERROR_cr_no_progress 204
#define ORC_ERROR_cr_inf_or_nan 205

/* lhs = AI*W*AI'*rhs for AI = [A I] given as CSC of all n+m columns; only the
 * first n columns are traversed. W == NULL: W = 1 on structurals, 0 on slacks.
 * *dot = rhs'lhs if dot != NULL. Reference src/normal_matrix.cc:45-126
 * (MATVECMETHOD 1). */
void orc_normal_apply(oint m, oint n, const oint* Ap, const oint* Ai,
                      const double* Ax, const double* W, const double* rhs,
                      double* lhs, double* dot);

/* diag[i] = W[n+i] + sum_j W[j]*a_ij^2 (W == NULL: sum_j a_ij^2).
 * Reference src/diagonal_precond.cc:28-46 without dense columns. */
void orc_diag_build(oint m, oint n, const oint* Ap, const oint* Ai,
                    const double* Ax, const double* W, double* diag);

/* lhs = rhs ./ diag, *dot = sum lhs[i]*rhs[i].
 * Reference src/diagonal_precond.cc:150-157. */
void orc_diag_apply(oint m, const double* diag, const double* rhs, double* lhs,
                    double* dot);

/* lhs += A*A'*rhs, or A*D^2*A'*rhs if D != NULL.
 * Reference src/sparse_matrix.cc:211-222. */
void orc_add_normal_product(oint nrow, oint ncol, const oint* Ap,
                            const oint* Ai, const double* Ax, const double* D,
                            const double* rhs, double* lhs);

/* ipx::MultiplyAdd (src/sparse_matrix.cc:194-209): trans 'N' lhs(nrow) +=
 * alpha*A*rhs(ncol), 't'/'T' lhs(ncol) += alpha*A'*rhs(nrow); the reference's
 * loops, so the sums are taken in the reference's order. */
void orc_multiply_add(oint nrow, oint ncol, const oint* Ap, const oint* Ai,
                      const double* Ax, const double* rhs, double alpha,
                      double* lhs, char trans);

/* In-place sparse triangular solve; returns nnz(x).
 * trans: 't'/'T' transposed. uplo: 'u'/'U' upper else lower. unitdiag != 0:
 * unit diagonal not stored; otherwise the diagonal is the LAST entry of each
 * column for upper, the FIRST for lower.
 * Reference src/sparse_matrix.cc:224-301. */
oint orc_triangular_solve(oint ncol, const oint* Ap, const oint* Ai,
                          const double* Ax, double* x, char trans, char uplo,
                          int unitdiag);

/* Reference src/sparse_matrix.cc:303-311. */
void orc_forward_solve(oint dim, const oint* Lp, const oint* Li,
                       const double* Lx, const oint* Up, const oint* Ui,
                       const double* Ux, double* x);
void orc_backward_solve(oint dim, const oint* Lp, const oint* Li,
                        const double* Lx, const oint* Up, const oint* Ui,
                        const double* Ux, double* x);

/* Basis-preconditioned operator C = I + inv(B) N N' inv(B') on prepared
 * factors. N has nrow = dim rows (already row-permuted and column-scaled), U is
 * column-scaled. work: dim doubles of scratch.
 * Reference src/splitted_normal_matrix.cc:90-117. */
typedef struct {
    oint dim;
    const oint *Lp, *Li; const double* Lx;
    const oint *Up, *Ui; const double* Ux;
    oint ncolN;
    const oint *Np, *Ni; const double* Nx;
    oint num_free;
    const oint* free_positions;
} orc_split;
void orc_split_apply(const orc_split* S, const double* rhs, double* lhs,
                     double* work, double* dot);

/* Operator selector for the CR drivers. */
typedef struct {
    int kind;                 /* 0: normal matrix, 1: split operator */
    oint m, n;                /* kind 0 */
    const oint *Ap, *Ai; const double* Ax; const double* W;
    const orc_split* split;   /* kind 1 */
    double* work;             /* kind 1: dim doubles */
} orc_operator;

/* Preconditioned CR with C = op, P = diag(1./diag).
 * lhs: initial iterate in, solution out. maxiter < 0 => m+100.
 * resnorm_hist (may be NULL): on return holds the residual norm tested at the
 * top of every loop pass (iter+1 entries if hist_cap allows).
 * Returns errflag (0, 201..205). Reference src/conjugate_residuals.cc:90-213. */
oint orc_pcr_solve(const orc_operator* op, oint m, const double* diag,
                   const double* rhs, double tol, const double* resscale,
                   oint maxiter, double* lhs, oint* iter, double* resnorm_hist,
                   oint hist_cap);

/* Unpreconditioned CR. Reference src/conjugate_residuals.cc:14-88. */
oint orc_cr_solve(const orc_operator* op, oint m, const double* rhs, double tol,
                  const double* resscale, oint maxiter, double* lhs, oint* iter,
                  double* resnorm_hist, oint hist_cap);

/* W[j] = 1/(zl/xl+zu/xu) with regularisation, resscale[i] = 1/sqrt(W[n+i]).
 * have_iterate == 0: W = 1. Reference src/kkt_solver_diag.cc:24-56. */
void orc_kktdiag_weights(oint m, oint n, int have_iterate, const double* xl,
                         const double* xu, const double* zl, const double* zu,
                         double mu, double* W, double* resscale);

/* rhs = -b + AI*(W.*a). Reference src/kkt_solver_diag.cc:90-92. */
void orc_kktdiag_rhs(oint m, oint n, const oint* Ap, const oint* Ai,
                     const double* Ax, const double* W, const double* a,
                     const double* b, double* rhs);

/* x from y. Reference src/kkt_solver_diag.cc:108-117. */
void orc_kktdiag_recover(oint m, oint n, const oint* Ap, const oint* Ai,
                         const double* Ax, const double* W, const double* a,
                         const double* b, const double* y, double* x);

/* Full KKTSolverDiag::_Solve: rhs assembly, y = 0, PCR, recovery.
 * Returns errflag; *iter = CR iterations. */
oint orc_kktdiag_solve(oint m, oint n, const oint* Ap, const oint* Ai,
                       const double* Ax, const double* W, const double* diag,
                       const double* resscale, const double* a, const double* b,
                       double tol, oint maxiter, double* x, double* y,
                       oint* iter);

#ifdef __cplusplus
}
#endif
#endif /* IPX_ORACLE_H_ */
