// GPU drop-in for ipx::ConjugateResiduals: defines the members declared in the
// UNMODIFIED reference header src/conjugate_residuals.h:21-71 and is linked
// instead of src/conjugate_residuals.cc.
//
// When the operators are the device-backed NormalMatrix / DiagonalPrecond /
// SplittedNormalMatrix (the only LinearOperator implementers in IPX), the whole
// loop runs device-resident through ipxgpu_pcr_solve / ipxgpu_cr_solve: the
// scalars never visit the host and Control::InterruptCheck is polled between
// batches of enqueued iterations. Any other LinearOperator pair (none exists
// in the reference tree; kept to honour the class contract, e.g. the
// dense-column preconditioner) is driven through Apply() from the host with
// the same sequence of tests.

#include "conjugate_residuals.h"

#include <algorithm>
#include <cmath>

#include "gpu_bridge.h"
#include "timer.h"
#include "utils.h"

namespace ipx {

using ipxb200::Check;
using ipxb200::OperatorKind;
using ipxb200::OperatorRecord;

namespace {

int64_t InterruptThunk(void* user) {
    return static_cast<const Control*>(user)->InterruptCheck();
}

// Device operator behind C, or nullptr.
OperatorRecord* DeviceOperator(LinearOperator& C) {
    OperatorRecord* rec = ipxb200::FindRecord(&C);
    if (!rec || !ipxb200::StillCurrent(*rec)) return nullptr;
    if (rec->kind != OperatorKind::kNormal && rec->kind != OperatorKind::kSplit) return nullptr;
    return rec;
}

void AddTimes(OperatorRecord* C, OperatorRecord* P, const ipxgpu_cr_result& res) {
    if (C->kind == OperatorKind::kNormal) {
        if (C->time) *C->time += res.time_op;
    } else {
        if (C->time_B) *C->time_B += res.time_B;
        if (C->time_Bt) *C->time_Bt += res.time_Bt;
        if (C->time_NNt) *C->time_NNt += res.time_NNt;
    }
    if (P && P->time) *P->time += res.time_pre;
}

double ScaledInfnorm(const Vector& v, const double* resscale) {
    if (!resscale) return Infnorm(v);
    double norm = 0.0;
    for (size_t i = 0; i < v.size(); i++) norm = std::max(norm, std::abs(resscale[i] * v[i]));
    return norm;
}

}  // namespace

ConjugateResiduals::ConjugateResiduals(const Control& control) : control_(control) {}

// Reference src/conjugate_residuals.cc:14-88.
void ConjugateResiduals::Solve(LinearOperator& C, const Vector& rhs, double tol,
                               const double* resscale, Int maxiter, Vector& lhs) {
    const Int m = rhs.size();
    Timer timer;
    errflag_ = 0;
    iter_ = 0;
    time_ = 0.0;

    if (OperatorRecord* dev = DeviceOperator(C)) {
        ipxgpu_cr_result res{};
        if (m > 0) {
            const int op = dev->kind == OperatorKind::kSplit ? 1 : 0;
            Check(ipxgpu_cr_solve(dev->ref.ctx, op, &rhs[0], tol, resscale, maxiter, &lhs[0], &res,
                                  InterruptThunk, const_cast<Control*>(&control_), nullptr, 0));
            AddTimes(dev, nullptr, res);
        }
        errflag_ = res.errflag;
        iter_ = res.iter;
        if (errflag_ == IPX_ERROR_cr_iter_limit)
            control_.Debug(3) << " CR method not converged in " << iter_ << " iterations."
                              << " residual = " << sci2(res.resnorm) << ','
                              << " tolerance = " << sci2(tol) << '\n';
        time_ = timer.Elapsed();
        return;
    }

    // Host-driven loop over Apply() for operators without a device twin.
    Vector residual(m), step(m), Cresidual(m), Cstep(m);
    double cdot = 0.0;
    if (maxiter < 0) maxiter = m + 100;
    if (Infnorm(lhs) == 0.0) {
        residual = rhs;
    } else {
        C.Apply(lhs, residual, nullptr);
        residual = rhs - residual;
    }
    C.Apply(residual, Cresidual, &cdot);
    step = residual;
    Cstep = Cresidual;
    for (;;) {
        const double resnorm = ScaledInfnorm(residual, resscale);
        if (resnorm <= tol) break;
        if (iter_ == maxiter) { errflag_ = IPX_ERROR_cr_iter_limit; break; }
        if (cdot <= 0.0) { errflag_ = IPX_ERROR_cr_matrix_not_posdef; break; }
        const double alpha = cdot / Dot(Cstep, Cstep);
        if (!std::isfinite(alpha)) { errflag_ = IPX_ERROR_cr_inf_or_nan; break; }
        lhs += alpha * step;
        residual -= alpha * Cstep;
        double cdotnew;
        C.Apply(residual, Cresidual, &cdotnew);
        const double beta = cdotnew / cdot;
        step = residual + beta * step;
        Cstep = Cresidual + beta * Cstep;
        cdot = cdotnew;
        iter_++;
        if ((errflag_ = control_.InterruptCheck()) != 0) break;
    }
    time_ = timer.Elapsed();
}

// Reference src/conjugate_residuals.cc:90-213.
void ConjugateResiduals::Solve(LinearOperator& C, LinearOperator& P, const Vector& rhs,
                               double tol, const double* resscale, Int maxiter, Vector& lhs) {
    const Int m = rhs.size();
    Timer timer;
    errflag_ = 0;
    iter_ = 0;
    time_ = 0.0;

    OperatorRecord* devC = DeviceOperator(C);
    OperatorRecord* devP = ipxb200::FindRecord(&P);
    const bool device_loop = devC && devC->kind == OperatorKind::kNormal && devP &&
                             devP->kind == OperatorKind::kDiagonal && !devP->host_part &&
                             ipxb200::StillCurrent(*devP) && devP->ref.ctx == devC->ref.ctx;
    if (device_loop) {
        ipxgpu_cr_result res{};
        if (m > 0) {
            Check(ipxgpu_pcr_solve(devC->ref.ctx, &rhs[0], tol, resscale, maxiter, &lhs[0], &res,
                                   InterruptThunk, const_cast<Control*>(&control_), nullptr, 0));
            AddTimes(devC, devP, res);
        }
        errflag_ = res.errflag;
        iter_ = res.iter;
        if (errflag_ == IPX_ERROR_cr_iter_limit)
            control_.Debug(3) << " PCR method not converged in " << iter_ << " iterations."
                              << " residual = " << sci2(res.resnorm) << ','
                              << " tolerance = " << sci2(tol) << '\n';
        else if (errflag_ == IPX_ERROR_cr_matrix_not_posdef)
            control_.Debug(3) << " matrix in PCR method not posdef.\n";
        else if (errflag_ == IPX_ERROR_cr_no_progress)
            control_.Debug(3) << " PCR method: preconditioned residual norm did not decrease.\n";
        time_ = timer.Elapsed();
        return;
    }

    // Host-driven loop (e.g. preconditioner with a dense-column part).
    Vector residual(m), sresidual(m), step(m), Csresidual(m), Cstep(m), PCstep(m);
    double cdot = 0.0, rho = 0.0;
    if (maxiter < 0) maxiter = m + 100;
    if (Infnorm(lhs) == 0.0) {
        residual = rhs;
    } else {
        C.Apply(lhs, residual, nullptr);
        residual = rhs - residual;
    }
    P.Apply(residual, sresidual, &rho);
    C.Apply(sresidual, Csresidual, &cdot);
    step = sresidual;
    Cstep = Csresidual;
    for (;;) {
        const double resnorm = ScaledInfnorm(residual, resscale);
        if (resnorm <= tol) break;
        if (iter_ == maxiter) { errflag_ = IPX_ERROR_cr_iter_limit; break; }
        if (cdot <= 0.0) { errflag_ = IPX_ERROR_cr_matrix_not_posdef; break; }
        double pdot;
        P.Apply(Cstep, PCstep, &pdot);
        if (pdot <= 0.0) { errflag_ = IPX_ERROR_cr_precond_not_posdef; break; }
        const double alpha = cdot / pdot;
        if (!std::isfinite(alpha)) { errflag_ = IPX_ERROR_cr_inf_or_nan; break; }
        lhs += alpha * step;
        residual -= alpha * Cstep;
        sresidual -= alpha * PCstep;
        double cdotnew;
        C.Apply(sresidual, Csresidual, &cdotnew);
        const double beta = cdotnew / cdot;
        step = sresidual + beta * step;
        Cstep = Csresidual + beta * Cstep;
        cdot = cdotnew;
        iter_++;
        if (iter_ % 5 == 0) {
            double rho_new;
            P.Apply(residual, sresidual, &rho_new);
            if (rho_new >= rho) { errflag_ = IPX_ERROR_cr_no_progress; break; }
            rho = rho_new;
        }
        if ((errflag_ = control_.InterruptCheck()) != 0) break;
    }
    time_ = timer.Elapsed();
}

Int ConjugateResiduals::errflag() const { return errflag_; }
Int ConjugateResiduals::iter() const { return iter_; }
double ConjugateResiduals::time() const { return time_; }

}  // namespace ipx
