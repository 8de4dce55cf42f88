// GPU drop-in for ipx::KKTSolverDiag: defines the members declared in the
// UNMODIFIED reference header src/kkt_solver_diag.h:23-47 and is linked instead
// of src/kkt_solver_diag.cc.
//
// Default route (no dense-column preconditioning in effect): the weight build,
// the diagonal, the right-hand side, the CR loop and the solution recovery all
// run on the device (ipxgpu_kktdiag_factorize / ipxgpu_kktdiag_solve); the host
// uploads the iterate's four vectors once per IPM iteration and (a, b) once
// per solve, and receives (x, y). W_ and resscale_ are mirrored back so the
// object's members keep their documented meaning.
// With dense columns and precond_dense_cols the preconditioner has a host part
// (diagonal_precond_gpu.cc); then the member operators are used and only their
// Apply()s run on the device.

#include "kkt_solver_diag.h"

#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "conjugate_residuals.h"
#include "gpu_bridge.h"

namespace ipx {

using ipxb200::Check;

namespace {
int64_t InterruptThunk(void* user) {
    return static_cast<const Control*>(user)->InterruptCheck();
}
}  // namespace

KKTSolverDiag::KKTSolverDiag(const Control& control, const Model& model)
    : control_(control), model_(model), normal_matrix_(model), precond_(model) {
    const Int m = model_.rows();
    const Int n = model_.cols();
    W_.resize(m + n);
    resscale_.resize(m);
}

void KKTSolverDiag::_Factorize(Iterate* pt, Info* info) {
    const Int m = model_.rows();
    const Int n = model_.cols();
    static const bool timing = std::getenv("IPXGPU_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[KKTSolverDiag::_Factorize] %-24s %8.2f ms\n", what,
                     1e3 * std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    iter_ = 0;
    factorized_ = false;
    const bool host_precond = control_.precond_dense_cols() && model_.num_dense_cols() > 0;

    if (!host_precond) {
        const ipxb200::ContextRef ref = ipxb200::ContextFor(model_);
        lap("ContextFor");
        if (pt) {
            Check(ipxgpu_kktdiag_factorize(ref.ctx, &pt->xl()[0], &pt->xu()[0], &pt->zl()[0],
                                           &pt->zu()[0], pt->mu(), &W_[0],
                                           m > 0 ? &resscale_[0] : nullptr));
        } else {
            Check(ipxgpu_kktdiag_factorize(ref.ctx, nullptr, nullptr, nullptr, nullptr, 0.0, &W_[0],
                                           m > 0 ? &resscale_[0] : nullptr));
        }
        // Weights and diagonal are already resident: the member operators only
        // register themselves (and fetch the diagonal).
        ipxb200::SetResidentHint(ref.ctx, &W_[0]);
        lap("ipxgpu_kktdiag_factorize");
    } else {
        // Host weight build, reference src/kkt_solver_diag.cc:24-56.
        if (pt) {
            const Vector &xl = pt->xl(), &xu = pt->xu(), &zl = pt->zl(), &zu = pt->zu();
            double regval = pt->mu();
            for (Int j = 0; j < n + m; j++) {
                const double g = zl[j] / xl[j] + zu[j] / xu[j];
                if (g != 0.0 && g < regval) regval = g;
                W_[j] = 1.0 / g;
            }
            for (Int j = 0; j < n + m; j++)
                if (std::isinf(W_[j])) W_[j] = 1.0 / regval;
        } else {
            W_ = 1.0;
        }
        for (Int i = 0; i < m; i++) resscale_[i] = 1.0 / std::sqrt(W_[n + i]);
    }
    normal_matrix_.Prepare(&W_[0]);
    lap("NormalMatrix::Prepare");
    precond_.Factorize(&W_[0], control_.precond_dense_cols(), info);
    lap("DiagonalPrecond::Factorize");
    if (info->errflag) return;
    factorized_ = true;
}

void KKTSolverDiag::_Solve(const Vector& a, const Vector& b, double tol, Vector& x, Vector& y,
                           Info* info) {
    const Int m = model_.rows();
    const Int n = model_.cols();
    assert(factorized_);
    const bool host_precond = control_.precond_dense_cols() && model_.num_dense_cols() > 0;
    const ipxb200::ContextRef ref = ipxb200::CurrentContext(model_);

    if (!host_precond && m > 0 && ref.ctx) {
        ipxgpu_cr_result res{};
        int rc = ipxgpu_kktdiag_solve(ref.ctx, &a[0], &b[0], tol, maxiter_, &x[0], &y[0], &res,
                                      InterruptThunk, const_cast<Control*>(&control_));
        if (rc == IPXGPU_ERR_STATE) {
            // The device context was rebuilt since Factorize (cache eviction).
            throw std::logic_error("KKTSolverDiag: device context lost; call Factorize again");
        }
        Check(rc);
        if (res.errflag == IPX_ERROR_cr_iter_limit)
            control_.Debug(3) << " PCR method not converged in " << res.iter << " iterations."
                              << " residual = " << sci2(res.resnorm) << ','
                              << " tolerance = " << sci2(tol) << '\n';
        info->errflag = res.errflag;
        info->kktiter1 += res.iter;
        info->time_cr1 += res.time;
        info->time_cr1_AAt += res.time_op;
        info->time_cr1_pre += res.time_pre;
        iter_ += res.iter;
        return;
    }

    // Member-operator route (reference src/kkt_solver_diag.cc:82-118).
    const SparseMatrix& AI = model_.AI();
    Vector rhs = -b;
    for (Int j = 0; j < n + m; j++) ScatterColumn(AI, j, W_[j] * a[j], rhs);
    y = 0.0;
    normal_matrix_.reset_time();
    precond_.reset_time();
    ConjugateResiduals cr(control_);
    cr.Solve(normal_matrix_, precond_, rhs, tol, &resscale_[0], maxiter_, y);
    info->errflag = cr.errflag();
    info->kktiter1 += cr.iter();
    info->time_cr1 += cr.time();
    info->time_cr1_AAt += normal_matrix_.time();
    info->time_cr1_pre += precond_.time();
    iter_ += cr.iter();
    for (Int i = 0; i < m; i++) x[n + i] = b[i];
    for (Int j = 0; j < n; j++) {
        x[j] = W_[j] * (a[j] - DotColumn(AI, j, y));
        for (Int p = AI.begin(j); p < AI.end(j); p++) x[n + AI.index(p)] -= x[j] * AI.value(p);
    }
}

}  // namespace ipx
