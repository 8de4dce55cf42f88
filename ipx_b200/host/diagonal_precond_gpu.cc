// GPU drop-in for ipx::DiagonalPrecond: defines the members declared in the
// UNMODIFIED reference header src/diagonal_precond.h:25-52 and is linked
// instead of src/diagonal_precond.cc.
//
// The diagonal E of AI*W*AI' is built on the device (one sweep over the CSR copy of A). With
// dense columns (src/model.cc:34-56, at most 1000) and precond_dense_cols, those columns are
// masked out of that sweep and the preconditioner becomes the Sherman-Morrison-Woodbury form
// inv(E) - inv(E) Ad inv(S) Ad' inv(E) (src/diagonal_precond.cc:48-102): the small Schur
// complement S = inv(Wd) + Ad' inv(E) Ad is assembled and Cholesky-factorized on the host
// once per Factorize (LAPACK dpotrf, as in the reference) and its factor is handed to the
// device, where every Apply - stand-alone or inside the CR loop - runs (ipxgpu_smw_load,
// csrc/smw.cuh). Nothing of Apply stays on the host.

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

namespace {

// Installs the dense-column part on the device: Ad' lives in Atdense (row i of Ad = column i
// of Atdense), the Cholesky factor in chol (nd x nd, lower).
void LoadDenseColumnPart(ipxgpu_ctx* ctx, const SparseMatrix& Atdense, const Vector& chol) {
    const Int nd = Atdense.rows();
    const SparseMatrix Ad = Transpose(Atdense);
    Check(ipxgpu_smw_load(ctx, nd, Ad.colptr(), Ad.rowidx(), Ad.values(), &chol[0]));
}

}  // namespace

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
    OperatorRecord& rec = ipxb200::RecordOf(this);
    rec.kind = OperatorKind::kDiagonal;
    rec.ref = ref;
    rec.model = &model_;
    rec.time = &time_;
    // Puts E (and the dense-column factor) of THIS object back on the device: after the context
    // was rebuilt, or after another preconditioner on the same model was factorized.
    rec.reprime = [this] {
        OperatorRecord& r = ipxb200::RecordOf(this);
        r.ref = ipxb200::ContextFor(model_);
        if (model_.rows() > 0) Check(ipxgpu_diag_set(r.ref.ctx, &diagonal_[0]));
        if (Atdense_.rows() > 0) LoadDenseColumnPart(r.ref.ctx, Atdense_, chol_factor_);
        else Check(ipxgpu_smw_clear(r.ref.ctx));
        ipxb200::ClaimState(r.ref.ctx, ipxb200::StateSlot::kDiagonal, this);
    };
    ipxb200::ClaimState(ref.ctx, ipxb200::StateSlot::kDiagonal, this);

    if (!smw) {
        // diag(AI*W*AI') over all columns, on the device (possibly already there).
        if (!ipxb200::ConsumeDiagonalHint(ref.ctx, W)) Check(ipxgpu_diag_factorize(ref.ctx, W, 0));
        else Check(ipxgpu_smw_clear(ref.ctx));
        if (m > 0) Check(ipxgpu_diag_get(ref.ctx, &diagonal_[0]));
        Atdense_.clear();
        chol_factor_.resize(0);
        work_.resize(0);
        factorized_ = true;
        return;
    }

    // E: the same sweep with the dense columns' weights zeroed on the device - they never
    // enter the sum, exactly as in the reference's loop (:28-36).
    std::vector<Int> dense;
    for (Int j = 0; j < n; j++)
        if (model_.IsDenseColumn(j)) dense.push_back(j);
    const Int nd = static_cast<Int>(dense.size());
    // (a resident full diagonal is not E, but resident weights spare the upload)
    const bool resident = ipxb200::ConsumeDiagonalHint(ref.ctx, W);
    Check(ipxgpu_diag_factorize_masked(ref.ctx, W, resident ? 1 : 0, nd, dense.data()));
    if (m > 0) Check(ipxgpu_diag_get(ref.ctx, &diagonal_[0]));

    // S = inv(Wd) + Ad' inv(E) Ad, lower triangle column by column: entry (l, k) is the
    // E-weighted inner product of dense columns l and k, gathered through the rows of Ad'.
    Atdense_ = Transpose(CopyColumns(AI, dense));
    chol_factor_.resize(nd * nd);
    chol_factor_ = 0.0;
    for (Int k = 0; k < nd; k++) {
        double* column = &chol_factor_[k * nd];
        const Int j = dense[k];
        for (Int p = AI.begin(j); p < AI.end(j); p++) {
            const Int i = AI.index(p);
            const double t = AI.value(p) / diagonal_[i];
            for (Int q = Atdense_.begin(i); q < Atdense_.end(i); q++)
                column[Atdense_.index(q)] += t * Atdense_.value(q);
        }
        column[k] += 1.0 / (W ? W[j] : 1.0);
    }
    if (Lapack_dpotrf('L', nd, &chol_factor_[0], nd) != 0) {
        info->errflag = IPX_ERROR_lapack_chol;
        return;
    }
    work_.resize(nd);
    LoadDenseColumnPart(ref.ctx, Atdense_, chol_factor_);
    factorized_ = true;
}

double DiagonalPrecond::time() const { return time_; }

void DiagonalPrecond::reset_time() { time_ = 0.0; }

void DiagonalPrecond::_Apply(const Vector& rhs, Vector& lhs, double* rhs_dot_lhs) {
    const Int m = model_.rows();
    Timer timer;
    assert(factorized_);
    assert((Int)lhs.size() == m);
    assert((Int)rhs.size() == m);

    OperatorRecord& rec = ipxb200::RecordOf(this);
    if (!ipxb200::StillCurrent(rec) ||
        !ipxb200::OwnsState(rec.ref.ctx, ipxb200::StateSlot::kDiagonal, this))
        rec.reprime();
    double dot = 0.0;
    if (m > 0) Check(ipxgpu_diag_apply(rec.ref.ctx, &rhs[0], &lhs[0], &dot));
    if (rhs_dot_lhs) *rhs_dot_lhs = dot;
    time_ += timer.Elapsed();
}

}  // namespace ipx
