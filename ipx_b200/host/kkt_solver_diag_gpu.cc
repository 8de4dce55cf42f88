// GPU drop-in for ipx::KKTSolverDiag: defines the members declared in the
// UNMODIFIED reference header src/kkt_solver_diag.h:23-47 and is linked instead
// of src/kkt_solver_diag.cc.
//
// The weight build, the diagonal, the right-hand side, the CR loop and the solution recovery
// all run on the device (ipxgpu_kktdiag_factorize / ipxgpu_kktdiag_solve); the host uploads
// the iterate's four vectors once per IPM iteration and (a, b) once per solve, and receives
// (x, y). W_ and resscale_ are mirrored back so the object's members keep their documented
// meaning. With dense columns and precond_dense_cols the preconditioner gets its
// Sherman-Morrison-Woodbury part in DiagonalPrecond::Factorize (diagonal_precond_gpu.cc); the
// solve is the same device call. There is no host route.

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

// IPXGPU_NGPUS > 1: the solver runs on a multi-GPU group (gpu_bridge.h). The dense-column
// preconditioner is a replicated small dense problem and keeps the single-GPU context.
bool UseGroup(const Control& control, const Model& model) {
    return ipxb200::GroupSize() > 1 && model.rows() > 0 &&
           !(control.precond_dense_cols() && model.num_dense_cols() > 0);
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
    if (UseGroup(control_, model_)) {
        // weights, diagonal and (in _Solve) the whole KKT solve on the group; the member
        // operators are not primed - nothing outside this class reaches them
        const ipxb200::ContextRef ref = ipxb200::GroupContextFor(model_);
        lap("GroupContextFor");
        Check(ipxgpu_kktdiag_factorize(ref.ctx, pt ? &pt->xl()[0] : nullptr,
                                       pt ? &pt->xu()[0] : nullptr, pt ? &pt->zl()[0] : nullptr,
                                       pt ? &pt->zu()[0] : nullptr, pt ? pt->mu() : 0.0, &W_[0],
                                       &resscale_[0]));
        lap("ipxgpu_kktdiag_factorize (group)");
        ipxb200::ClaimState(ref.ctx, ipxb200::StateSlot::kKktDiag, this);
        factorized_ = true;
        return;
    }
    {
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
        ipxb200::ClaimState(ref.ctx, ipxb200::StateSlot::kKktDiag, this);
        lap("ipxgpu_kktdiag_factorize");
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
    const ipxb200::ContextRef ref = UseGroup(control_, model_)
                                        ? ipxb200::CurrentGroupContext(model_)
                                        : ipxb200::CurrentContext(model_);
    if (m == 0) {
        // no constraints: x = W a (reference :108-117 with empty sums)
        for (Int j = 0; j < n; j++) x[j] = W_[j] * a[j];
        return;
    }
    if (!ref.ctx) throw std::logic_error("KKTSolverDiag: no device context; call Factorize first");
    if (!ipxb200::OwnsState(ref.ctx, ipxb200::StateSlot::kKktDiag, this))
        throw std::logic_error("KKTSolverDiag: another solver was factorized on this model since; "
                               "call Factorize again");
    // weights and preconditioner on the device belong to the member operators
    if (!UseGroup(control_, model_)) {
        if (ipxb200::OperatorRecord* c = ipxb200::FindRecord(&normal_matrix_))
            ipxb200::EnsurePrimed(*c, &normal_matrix_);
        if (ipxb200::OperatorRecord* p = ipxb200::FindRecord(&precond_))
            ipxb200::EnsurePrimed(*p, &precond_);
    }

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
    else if (res.errflag == IPX_ERROR_cr_matrix_not_posdef)
        control_.Debug(3) << " matrix in PCR method not posdef.\n";
    else if (res.errflag == IPX_ERROR_cr_no_progress)
        control_.Debug(3) << " PCR method: preconditioned residual norm did not decrease.\n";
    info->errflag = res.errflag;
    info->kktiter1 += res.iter;
    info->time_cr1 += res.time;
    info->time_cr1_AAt += res.time_op;
    info->time_cr1_pre += res.time_pre;
    iter_ += res.iter;
}

}  // namespace ipx
