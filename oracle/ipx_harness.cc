// C entry points over IPX's own operator / KKT-solver classes.
//
// TEST INFRASTRUCTURE. This TU is compiled against the UNMODIFIED reference headers and
// linked into two shared libraries that export the same symbols:
//   oracle/_ref/libipx_ref.so          with the reference CPU objects (the parity oracle
//                                      and the CPU baseline), and
//   oracle/_ref/libipx_gpu_harness.so  alone, against ipx_b200/_build/libipx_gpu.so - the
//                                      same reference objects EXCEPT the seven substituted TUs,
//                                      which are replaced by the GPU drop-ins of
//                                      ipx_b200/host. The product library holds no test code.
// Tests drive both through identical calls, so a parity test reads like a test
// of the reference's own classes (NormalMatrix::Apply, DiagonalPrecond::
// Factorize/Apply, ConjugateResiduals::Solve, KKTSolverDiag, Basis,
// SplittedNormalMatrix, KKTSolverBasis). The unchanged public API
// (include/ipx_c.h) is exported by both libraries as well.

#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "basis.h"
#include "conjugate_residuals.h"
#include "control.h"
#include "diagonal_precond.h"
#include "iterate.h"
#include "kkt_solver_basis.h"
#include "kkt_solver_diag.h"
#include "maxvolume.h"
#include "model.h"
#include "normal_matrix.h"
#include "presolver.h"
#include "sparse_matrix.h"
#include "splitted_normal_matrix.h"
#include "user_model.h"
#include "utils.h"

using namespace ipx;

namespace {

struct Harness {
    Control control;
    UserModel user_model;
    Model model;
    std::unique_ptr<Presolver> presolver;
    std::unique_ptr<NormalMatrix> normal;
    std::unique_ptr<DiagonalPrecond> precond;
    std::unique_ptr<KKTSolverDiag> kkt_diag;
    std::unique_ptr<Iterate> iterate;
    std::unique_ptr<Basis> basis;
    std::unique_ptr<SplittedNormalMatrix> split;
    std::unique_ptr<KKTSolverBasis> kkt_basis;
    Vector W;  // weights captured by NormalMatrix::Prepare (pointer semantics)
    Info info;
};

Vector ToVector(const double* p, Int n) {
    Vector v(n);
    if (n > 0) std::memcpy(&v[0], p, n * sizeof(double));
    return v;
}

void FromVector(const Vector& v, double* p) {
    if (v.size() > 0) std::memcpy(p, &v[0], v.size() * sizeof(double));
}

void FillInfoOut(const Info& info, double* out) {
    // [errflag, kktiter1, kktiter2, time_cr1, time_cr1_AAt, time_cr1_pre,
    //  time_cr2, time_cr2_NNt, time_cr2_B, time_cr2_Bt, time_kkt_factorize,
    //  time_kkt_solve, updates_ipm, primal_dropped, dual_dropped, time_maxvol]
    if (!out) return;
    out[0] = info.errflag;
    out[1] = info.kktiter1;
    out[2] = info.kktiter2;
    out[3] = info.time_cr1;
    out[4] = info.time_cr1_AAt;
    out[5] = info.time_cr1_pre;
    out[6] = info.time_cr2;
    out[7] = info.time_cr2_NNt;
    out[8] = info.time_cr2_B;
    out[9] = info.time_cr2_Bt;
    out[10] = info.time_kkt_factorize;
    out[11] = info.time_kkt_solve;
    out[12] = info.updates_ipm;
    out[13] = info.primal_dropped;
    out[14] = info.dual_dropped;
    out[15] = info.time_maxvol;
}

}  // namespace

extern "C" {

// Builds Control -> UserModel -> Presolver -> Model (the only way to populate
// a Model, src/model.h:84). @params may be NULL (defaults with display=0).
void* ipxh_create(ipxint num_constr, ipxint num_var, const ipxint* Ap,
                  const ipxint* Ai, const double* Ax, const double* rhs,
                  const char* constr_type, const double* obj, const double* lb,
                  const double* ub, const ipx_parameters* params,
                  ipxint* errflag) {
    std::unique_ptr<Harness> h(new Harness);
    ipx_parameters p;
    if (params) p = *params;
    else p.display = 0;
    h->control.parameters(p);
    Int err = h->user_model.Load(h->control, num_constr, num_var, Ap, Ai, Ax,
                                 rhs, constr_type, obj, lb, ub);
    if (errflag) *errflag = err;
    if (err) return nullptr;
    h->presolver.reset(new Presolver(h->user_model, h->model));
    err = h->presolver->PresolveModel(h->control);
    if (errflag) *errflag = err;
    if (err) return nullptr;
    return h.release();
}

void ipxh_free(void* self) { delete static_cast<Harness*>(self); }

void ipxh_dims(void* self, ipxint* m, ipxint* n, ipxint* nnz) {
    Harness* h = static_cast<Harness*>(self);
    *m = h->model.rows();
    *n = h->model.cols();
    *nnz = h->model.AI().entries();
}

void ipxh_get_AI(void* self, ipxint* AIp, ipxint* AIi, double* AIx) {
    const SparseMatrix& AI = static_cast<Harness*>(self)->model.AI();
    std::memcpy(AIp, AI.colptr(), (AI.cols() + 1) * sizeof(ipxint));
    std::memcpy(AIi, AI.rowidx(), AI.entries() * sizeof(ipxint));
    std::memcpy(AIx, AI.values(), AI.entries() * sizeof(double));
}

void ipxh_get_model_vectors(void* self, double* b, double* c, double* lb,
                            double* ub) {
    const Model& model = static_cast<Harness*>(self)->model;
    if (b) FromVector(model.b(), b);
    if (c) FromVector(model.c(), c);
    if (lb) FromVector(model.lb(), lb);
    if (ub) FromVector(model.ub(), ub);
}

// ipx::MultiplyAdd on the model's AI (src/sparse_matrix.cc:194-209): trans 'N'
// lhs(m) += alpha*AI*rhs(n+m), 'T' lhs(n+m) += alpha*AI'*rhs(m).
void ipxh_multiply_add_AI(void* self, const double* rhs, double alpha,
                          double* lhs, char trans) {
    const Model& model = static_cast<Harness*>(self)->model;
    const Int m = model.rows(), nm = model.rows() + model.cols();
    const bool t = trans == 't' || trans == 'T';
    Vector r = ToVector(rhs, t ? m : nm), l = ToVector(lhs, t ? nm : m);
    MultiplyAdd(model.AI(), r, alpha, l, trans);
    FromVector(l, lhs);
}

// ---- NormalMatrix (src/normal_matrix.h) ----

void ipxh_normal_prepare(void* self, const double* W) {
    Harness* h = static_cast<Harness*>(self);
    const Int nm = h->model.rows() + h->model.cols();
    if (!h->normal) h->normal.reset(new NormalMatrix(h->model));
    if (W) {
        h->W = ToVector(W, nm);
        h->normal->Prepare(&h->W[0]);
    } else {
        h->normal->Prepare(nullptr);
    }
}

void ipxh_normal_apply(void* self, const double* rhs, double* lhs,
                       double* rhs_dot_lhs) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l(m);
    h->normal->Apply(r, l, rhs_dot_lhs);
    FromVector(l, lhs);
}

// Repeats Apply @reps times on host-resident Vectors; returns the class's own
// time() accumulator (src/normal_matrix.cc:57,125) divided by reps.
double ipxh_normal_apply_timed(void* self, const double* rhs, double* lhs,
                               ipxint reps) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l(m);
    double dot;
    h->normal->reset_time();
    for (Int k = 0; k < reps; k++) h->normal->Apply(r, l, &dot);
    FromVector(l, lhs);
    return h->normal->time() / (reps > 0 ? reps : 1);
}

// ---- DiagonalPrecond (src/diagonal_precond.h) ----

ipxint ipxh_diag_factorize(void* self, const double* W,
                           ipxint precond_dense_cols) {
    Harness* h = static_cast<Harness*>(self);
    if (!h->precond) h->precond.reset(new DiagonalPrecond(h->model));
    h->info.errflag = 0;
    h->precond->Factorize(W, precond_dense_cols != 0, &h->info);
    return h->info.errflag;
}

void ipxh_diag_apply(void* self, const double* rhs, double* lhs,
                     double* rhs_dot_lhs) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l(m);
    h->precond->Apply(r, l, rhs_dot_lhs);
    FromVector(l, lhs);
}

// ---- ConjugateResiduals (src/conjugate_residuals.h) ----

// Preconditioned CR with C = the prepared NormalMatrix, P = the factorized
// DiagonalPrecond. @out = [errflag, iter, time].
void ipxh_pcr_solve(void* self, const double* rhs, double tol,
                    const double* resscale, ipxint maxiter, double* lhs,
                    double* out) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l = ToVector(lhs, m);
    ConjugateResiduals cr(h->control);
    cr.Solve(*h->normal, *h->precond, r, tol, resscale, maxiter, l);
    FromVector(l, lhs);
    out[0] = cr.errflag();
    out[1] = cr.iter();
    out[2] = cr.time();
}

// Unpreconditioned CR with C = the prepared NormalMatrix.
void ipxh_cr_solve_normal(void* self, const double* rhs, double tol,
                          const double* resscale, ipxint maxiter, double* lhs,
                          double* out) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l = ToVector(lhs, m);
    ConjugateResiduals cr(h->control);
    cr.Solve(*h->normal, r, tol, resscale, maxiter, l);
    FromVector(l, lhs);
    out[0] = cr.errflag();
    out[1] = cr.iter();
    out[2] = cr.time();
}

// ---- Iterate ----

void ipxh_iterate_set(void* self, const double* x, const double* xl,
                      const double* xu, const double* y, const double* zl,
                      const double* zu) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows(), n = h->model.cols();
    h->iterate.reset(new Iterate(h->model));
    h->iterate->Initialize(ToVector(x, n + m), ToVector(xl, n + m),
                           ToVector(xu, n + m), ToVector(y, m),
                           ToVector(zl, n + m), ToVector(zu, n + m));
}

// ---- KKTSolverDiag (src/kkt_solver_diag.h) ----

void ipxh_kktdiag_maxiter(void* self, ipxint maxiter) {
    Harness* h = static_cast<Harness*>(self);
    if (!h->kkt_diag)
        h->kkt_diag.reset(new KKTSolverDiag(h->control, h->model));
    h->kkt_diag->maxiter(maxiter);
}

// Factorize from the Iterate set by ipxh_iterate_set, or with G = I when
// @use_iterate == 0 (src/kkt_solver_diag.cc:50-52).
ipxint ipxh_kktdiag_factorize(void* self, ipxint use_iterate) {
    Harness* h = static_cast<Harness*>(self);
    if (!h->kkt_diag)
        h->kkt_diag.reset(new KKTSolverDiag(h->control, h->model));
    h->info = Info();
    h->kkt_diag->Factorize(use_iterate ? h->iterate.get() : nullptr, &h->info);
    return h->info.errflag;
}

ipxint ipxh_kktdiag_solve(void* self, const double* a, const double* b,
                          double tol, double* x, double* y, double* info_out) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows(), n = h->model.cols();
    Vector va = ToVector(a, n + m), vb = ToVector(b, m), vx(n + m), vy(m);
    h->kkt_diag->Solve(va, vb, tol, vx, vy, &h->info);
    FromVector(vx, x);
    FromVector(vy, y);
    FillInfoOut(h->info, info_out);
    return h->info.errflag;
}

ipxint ipxh_kktdiag_iter(void* self) {
    return static_cast<Harness*>(self)->kkt_diag->iter();
}

// ---- Basis / SplittedNormalMatrix / KKTSolverBasis ----

static Basis& GetBasis(Harness* h) {
    if (!h->basis) h->basis.reset(new Basis(h->control, h->model));
    return *h->basis;
}

ipxint ipxh_basis_load(void* self, const int* basic_status) {
    Harness* h = static_cast<Harness*>(self);
    return GetBasis(h).Load(basic_status);
}

ipxint ipxh_basis_from_weights(void* self, const double* colweights) {
    Harness* h = static_cast<Harness*>(self);
    h->info = Info();
    GetBasis(h).ConstructBasisFromWeights(colweights, &h->info);
    return h->info.errflag;
}

void ipxh_basis_get(void* self, ipxint* basis, int* status) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows(), n = h->model.cols();
    Basis& B = GetBasis(h);
    if (basis)
        for (Int p = 0; p < m; p++) basis[p] = B[p];
    if (status)
        for (Int j = 0; j < n + m; j++) status[j] = B.StatusOf(j);
}

void ipxh_basis_free_variable(void* self, ipxint j) {
    GetBasis(static_cast<Harness*>(self)).FreeBasicVariable(j);
}

void ipxh_basis_fix_variable(void* self, ipxint j) {
    GetBasis(static_cast<Harness*>(self)).FixNonbasicVariable(j);
}

// The factors can only be exported from a fresh factorization; after updates
// (crash basis repairs) refactorize first, as KKTSolverBasis::_Factorize does
// (src/kkt_solver_basis.cc:59-63).
static Basis& FreshBasis(Harness* h) {
    Basis& B = GetBasis(h);
    if (!B.FactorizationIsFresh()) B.Factorize();
    return B;
}

// Returns nnz(L), nnz(U) of the fresh factorization (B[rowperm,colperm] =
// (L+I)U, src/basis.h:109-118).
void ipxh_basis_lu_sizes(void* self, ipxint* lnz, ipxint* unz) {
    Harness* h = static_cast<Harness*>(self);
    SparseMatrix L, U;
    FreshBasis(h).GetLuFactors(&L, &U, nullptr, nullptr);
    *lnz = L.entries();
    *unz = U.entries();
}

void ipxh_basis_lu(void* self, ipxint* Lp, ipxint* Li, double* Lx, ipxint* Up,
                   ipxint* Ui, double* Ux, ipxint* rowperm, ipxint* colperm) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    SparseMatrix L, U;
    FreshBasis(h).GetLuFactors(&L, &U, rowperm, colperm);
    std::memcpy(Lp, L.colptr(), (m + 1) * sizeof(ipxint));
    std::memcpy(Li, L.rowidx(), L.entries() * sizeof(ipxint));
    std::memcpy(Lx, L.values(), L.entries() * sizeof(double));
    std::memcpy(Up, U.colptr(), (m + 1) * sizeof(ipxint));
    std::memcpy(Ui, U.rowidx(), U.entries() * sizeof(ipxint));
    std::memcpy(Ux, U.values(), U.entries() * sizeof(double));
}

void ipxh_basis_solve_dense(void* self, const double* rhs, double* lhs,
                            char trans) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l(m);
    GetBasis(h).SolveDense(r, l, trans);
    FromVector(l, lhs);
}

void ipxh_split_prepare(void* self, const double* colscale) {
    Harness* h = static_cast<Harness*>(self);
    if (!h->split) h->split.reset(new SplittedNormalMatrix(h->model));
    h->split->Prepare(FreshBasis(h), colscale);
}

void ipxh_split_colperm(void* self, ipxint* colperm) {
    Harness* h = static_cast<Harness*>(self);
    std::memcpy(colperm, h->split->colperm(), h->model.rows() * sizeof(ipxint));
}

void ipxh_split_apply(void* self, const double* rhs, double* lhs,
                      double* rhs_dot_lhs) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l(m);
    h->split->Apply(r, l, rhs_dot_lhs);
    FromVector(l, lhs);
}

// @times = [time_B, time_Bt, time_NNt] per apply, averaged over @reps.
void ipxh_split_apply_timed(void* self, const double* rhs, double* lhs,
                            ipxint reps, double* times) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l(m);
    double dot;
    h->split->reset_time();
    for (Int k = 0; k < reps; k++) h->split->Apply(r, l, &dot);
    FromVector(l, lhs);
    const double d = reps > 0 ? reps : 1;
    times[0] = h->split->time_B() / d;
    times[1] = h->split->time_Bt() / d;
    times[2] = h->split->time_NNt() / d;
}

// Unpreconditioned CR with C = the prepared SplittedNormalMatrix
// (the call of src/kkt_solver_basis.cc:150). @out = [errflag, iter, time].
void ipxh_cr_solve_split(void* self, const double* rhs, double tol,
                         ipxint maxiter, double* lhs, double* out) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows();
    Vector r = ToVector(rhs, m), l = ToVector(lhs, m);
    ConjugateResiduals cr(h->control);
    cr.Solve(*h->split, r, tol, nullptr, maxiter, l);
    FromVector(l, lhs);
    out[0] = cr.errflag();
    out[1] = cr.iter();
    out[2] = cr.time();
}

void ipxh_kktbasis_maxiter(void* self, ipxint maxiter) {
    Harness* h = static_cast<Harness*>(self);
    if (!h->kkt_basis)
        h->kkt_basis.reset(new KKTSolverBasis(h->control, GetBasis(h)));
    h->kkt_basis->maxiter(maxiter);
}

ipxint ipxh_kktbasis_factorize(void* self, double* info_out) {
    Harness* h = static_cast<Harness*>(self);
    if (!h->kkt_basis)
        h->kkt_basis.reset(new KKTSolverBasis(h->control, GetBasis(h)));
    h->info = Info();
    h->kkt_basis->Factorize(h->iterate.get(), &h->info);
    FillInfoOut(h->info, info_out);
    return h->info.errflag;
}

ipxint ipxh_kktbasis_solve(void* self, const double* a, const double* b,
                           double tol, double* x, double* y, double* info_out) {
    Harness* h = static_cast<Harness*>(self);
    const Int m = h->model.rows(), n = h->model.cols();
    Vector va = ToVector(a, n + m), vb = ToVector(b, m), vx(n + m), vy(m);
    h->kkt_basis->Solve(va, vb, tol, vx, vy, &h->info);
    FromVector(vx, x);
    FromVector(vy, y);
    FillInfoOut(h->info, info_out);
    return h->info.errflag;
}

// Maxvolume on the current basis (src/maxvolume.h:14-15). heuristic != 0: RunHeuristic, else
// RunSequential. colscale: n+m factors or NULL. out = [errflag, updates, skipped, passes,
// slices, volinc, time].
ipxint ipxh_maxvolume(void* self, const double* colscale, ipxint heuristic, double* out) {
    Harness* h = static_cast<Harness*>(self);
    Maxvolume maxvol(h->control);
    Basis& B = GetBasis(h);
    const Int errflag = heuristic ? maxvol.RunHeuristic(colscale, B)
                                  : maxvol.RunSequential(colscale, B);
    out[0] = errflag;
    out[1] = maxvol.updates();
    out[2] = maxvol.skipped();
    out[3] = maxvol.passes();
    out[4] = maxvol.slices();
    out[5] = maxvol.volinc();
    out[6] = maxvol.time();
    return errflag;
}

// ---- sparse kernels (src/sparse_matrix.h) on caller-supplied CSC arrays ----

static SparseMatrix MakeMatrix(ipxint nrow, ipxint ncol, const ipxint* Ap,
                               const ipxint* Ai, const double* Ax) {
    SparseMatrix A(nrow, ncol, Ap[ncol]);
    std::memcpy(A.colptr(), Ap, (ncol + 1) * sizeof(ipxint));
    std::memcpy(A.rowidx(), Ai, Ap[ncol] * sizeof(ipxint));
    std::memcpy(A.values(), Ax, Ap[ncol] * sizeof(double));
    return A;
}

ipxint ipxh_triangular_solve(ipxint dim, const ipxint* Ap, const ipxint* Ai,
                             const double* Ax, double* x, char trans,
                             char uplo, int unitdiag) {
    SparseMatrix A = MakeMatrix(dim, dim, Ap, Ai, Ax);
    Vector v = ToVector(x, dim);
    const char up[2] = {uplo, 0};
    Int nz = TriangularSolve(A, v, trans, up, unitdiag);
    FromVector(v, x);
    return nz;
}

void ipxh_add_normal_product(ipxint nrow, ipxint ncol, const ipxint* Ap,
                             const ipxint* Ai, const double* Ax,
                             const double* D, const double* rhs, double* lhs) {
    SparseMatrix A = MakeMatrix(nrow, ncol, Ap, Ai, Ax);
    Vector r = ToVector(rhs, nrow), l = ToVector(lhs, nrow);
    AddNormalProduct(A, D, r, l);
    FromVector(l, lhs);
}

}  // extern "C"
