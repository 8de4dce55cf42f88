// GPU drop-in for ipx::DiagonalPrecond: defines the members declared in the
// UNMODIFIED reference header src/diagonal_precond.h:25-52 and is linked
// instead of src/diagonal_precond.cc.
//
// The diagonal of AI*W*AI' is built on the device (one sweep over the CSR copy
// of A). The optional dense-column Sherman-Morrison-Woodbury part
// (src/diagonal_precond.cc:48-102,133-149; at most 1000 columns,
// src/model.cc:34-56) is a small dense problem that stays on the host, as in
// the reference; it is not exercised by the benchmark configurations.

#include "diagonal_precond.h"

#include <cassert>
#include <cmath>
#include <vector>

#include "gpu_bridge.h"
#include "lapack.h"
#include "timer.h"

namespace ipx {

using ipxb200::Check;
using ipxb200::OperatorKind;
using ipxb200::OperatorRecord;

DiagonalPrecond::DiagonalPrecond(const Model& model) : model_(model) {
    diagonal_.resize(model_.rows());
    ipxb200::Forget(this);
}

void DiagonalPrecond::Factorize(const double* W, bool precond_dense_cols, Info* info) {
    const Int m = model_.rows();
    const Int n = model_.cols();
    const SparseMatrix& AI = model_.AI();
    const bool smw = precond_dense_cols && model_.num_dense_cols() > 0;
    factorized_ = false;

    const ipxb200::ContextRef ref = ipxb200::ContextFor(model_);
    // diag(AI*W*AI') over ALL columns, on the device.
    if (!ipxb200::ConsumeDiagonalHint(ref.ctx, W)) Check(ipxgpu_diag_factorize(ref.ctx, W, 0));
    if (m > 0) Check(ipxgpu_diag_get(ref.ctx, &diagonal_[0]));

    std::vector<Int> dense;
    if (smw) {
        // The preconditioner's diagonal part E excludes the dense columns:
        // take their contribution out again (few columns, host).
        for (Int j = 0; j < n; j++)
            if (model_.IsDenseColumn(j)) dense.push_back(j);
        for (Int j : dense) {
            const double w = W ? W[j] : 1.0;
            for (Int p = AI.begin(j); p < AI.end(j); p++)
                diagonal_[AI.index(p)] -= AI.value(p) * w * AI.value(p);
        }
        if (m > 0) Check(ipxgpu_diag_set(ref.ctx, &diagonal_[0]));

        // inv(P) = inv(E) - inv(E) Ad inv(S) Ad' inv(E) with the Schur
        // complement S = inv(Wd) + Ad' inv(E) Ad.
        const Int nd = static_cast<Int>(dense.size());
        Atdense_ = Transpose(CopyColumns(AI, dense));
        chol_factor_.resize(nd * nd);
        chol_factor_ = 0.0;
        for (Int k = 0; k < nd; k++) {
            const Int j = dense[k];
            double* Scol = &chol_factor_[k * nd];
            for (Int p = AI.begin(j); p < AI.end(j); p++) {
                const Int i = AI.index(p);
                const double scaled = AI.value(p) / diagonal_[i];
                for (Int q = Atdense_.begin(i); q < Atdense_.end(i); q++)
                    Scol[Atdense_.index(q)] += scaled * Atdense_.value(q);
            }
            Scol[k] += 1.0 / (W ? W[j] : 1.0);
        }
        if (Lapack_dpotrf('L', nd, &chol_factor_[0], nd) != 0) {
            info->errflag = IPX_ERROR_lapack_chol;
            return;
        }
        work_.resize(nd);
    } else {
        Atdense_.clear();
        chol_factor_.resize(0);
        work_.resize(0);
    }

    OperatorRecord& rec = ipxb200::RecordOf(this);
    rec.kind = OperatorKind::kDiagonal;
    rec.ref = ref;
    rec.model = &model_;
    rec.host_part = smw;
    rec.time = &time_;
    factorized_ = true;
}

double DiagonalPrecond::time() const { return time_; }

void DiagonalPrecond::reset_time() { time_ = 0.0; }

void DiagonalPrecond::_Apply(const Vector& rhs, Vector& lhs, double* rhs_dot_lhs) {
    const Int m = model_.rows();
    const Int nd = Atdense_.rows();
    Timer timer;
    assert(factorized_);
    assert((Int)lhs.size() == m);
    assert((Int)rhs.size() == m);

    if (nd == 0) {
        OperatorRecord& rec = ipxb200::RecordOf(this);
        if (!ipxb200::StillCurrent(rec)) {  // context was rebuilt: reinstall
            rec.ref = ipxb200::ContextFor(model_);
            if (m > 0) Check(ipxgpu_diag_set(rec.ref.ctx, &diagonal_[0]));
        }
        double dot = 0.0;
        if (m > 0) Check(ipxgpu_diag_apply(rec.ref.ctx, &rhs[0], &lhs[0], &dot));
        if (rhs_dot_lhs) *rhs_dot_lhs = dot;
    } else {
        // Dense-column branch (host): lhs = inv(E) (rhs - Ad inv(S) Ad' inv(E) rhs).
        work_ = 0.0;
        for (Int i = 0; i < m; i++) ScatterColumn(Atdense_, i, rhs[i] / diagonal_[i], work_);
        Int err = Lapack_dpotrs('L', nd, 1, &chol_factor_[0], nd, &work_[0], nd);
        assert(err == 0);
        (void)err;
        double dot = 0.0;
        for (Int i = 0; i < m; i++) {
            lhs[i] = (rhs[i] - DotColumn(Atdense_, i, work_)) / diagonal_[i];
            dot += lhs[i] * rhs[i];
        }
        if (rhs_dot_lhs) *rhs_dot_lhs = dot;
    }
    time_ += timer.Elapsed();
}

}  // namespace ipx
