// GPU drop-in for ipx::ConjugateResiduals: defines the members declared in the
// UNMODIFIED reference header src/conjugate_residuals.h:21-71 and is linked
// instead of src/conjugate_residuals.cc.
//
// The operators are the device-backed NormalMatrix / DiagonalPrecond /
// SplittedNormalMatrix (the only LinearOperator implementers in IPX): the whole
// loop runs device-resident through ipxgpu_pcr_solve / ipxgpu_cr_solve, the
// scalars never visit the host and Control::InterruptCheck is polled while the
// device iterates. There is no host loop: a LinearOperator without a device twin
// (none exists in the reference tree) is refused with std::logic_error, which
// LpSolver::Solve reports as IPX_STATUS_internal_error (src/lp_solver.cc:98-105).

#include "conjugate_residuals.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

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
        ipxb200::EnsurePrimed(*dev, &C);
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

    throw std::logic_error("ConjugateResiduals: operator has no device twin (expected NormalMatrix "
                           "or SplittedNormalMatrix, prepared on the current device context)");
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
                             devP->kind == OperatorKind::kDiagonal &&
                             ipxb200::StillCurrent(*devP) && devP->ref.ctx == devC->ref.ctx;
    if (device_loop) {
        ipxgpu_cr_result res{};
        ipxb200::EnsurePrimed(*devC, &C);
        ipxb200::EnsurePrimed(*devP, &P);
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

    throw std::logic_error("ConjugateResiduals: operators have no device twins (expected "
                           "NormalMatrix and DiagonalPrecond of one model, prepared on the "
                           "current device context)");
}

Int ConjugateResiduals::errflag() const { return errflag_; }
Int ConjugateResiduals::iter() const { return iter_; }
double ConjugateResiduals::time() const { return time_; }

}  // namespace ipx
