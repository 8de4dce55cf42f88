// GPU drop-in for ipx::SplittedNormalMatrix: defines the members declared in
// the UNMODIFIED reference header src/splitted_normal_matrix.h:26-60 and is
// linked instead of src/splitted_normal_matrix.cc.
//
// Prepare pulls L, U and the permutations from the host Basis exactly like the
// reference (src/splitted_normal_matrix.cc:26-39) and hands them to the device,
// where the four triangular solves are level-scheduled. N is NOT materialised:
// N*N' is applied through the device-resident AI with the squared column
// scales of the NONBASIC variables (0 for every other column) and the row
// permutation, which is the same operator as scaling and permuting a copy of
// the nonbasic columns (:42-55).

#include "splitted_normal_matrix.h"

#include <cassert>
#include <cmath>
#include <stdexcept>

#include "gpu_bridge.h"
#include "timer.h"
#include "utils.h"

namespace ipx {

using ipxb200::Check;
using ipxb200::OperatorKind;
using ipxb200::OperatorRecord;

SplittedNormalMatrix::SplittedNormalMatrix(const Model& model) : model_(model) {
    const Int m = model_.rows();
    colperm_.resize(m);
    rowperm_inv_.resize(m);
    work_.resize(m);
    ipxb200::Forget(this);
}

void SplittedNormalMatrix::Prepare(const Basis& basis, const double* colscale) {
    const Int m = model_.rows();
    const Int n = model_.cols();
    assert(colscale);
    prepared_ = false;
    N_.clear();

    basis.GetLuFactors(&L_, &U_, rowperm_inv_.data(), colperm_.data());
    rowperm_inv_ = InversePerm(rowperm_inv_);

    // Columns of U that belong to BASIC (not BASIC_FREE) variables carry the
    // interior-point scaling; free positions become unit rows/columns of C.
    free_positions_.clear();
    for (Int k = 0; k < m; k++) {
        const Int j = basis[colperm_[k]];
        if (basis.StatusOf(j) == Basis::BASIC) {
            const double d = colscale[j];
            assert(std::isfinite(d) && d > 0.0);
            ScaleColumn(U_, k, d);
        } else if (basis.StatusOf(j) == Basis::BASIC_FREE) {
            free_positions_.push_back(k);
        }
    }

    // Scale factors of the NONBASIC columns; every other column (BASIC,
    // BASIC_FREE, NONBASIC_FIXED) is masked with 0 - their colscale may be
    // 0 or infinite (src/iterate.cc:183-198) and must never be multiplied in.
    Vector nonbasic_scale(0.0, n + m);
    for (Int j = 0; j < n + m; j++) {
        if (basis.StatusOf(j) == Basis::NONBASIC) {
            assert(std::isfinite(colscale[j]));
            nonbasic_scale[j] = colscale[j];
        }
    }

    const ipxb200::ContextRef ref = ipxb200::ContextFor(model_);
    Check(ipxgpu_lu_load(ref.ctx, m, L_.colptr(), L_.rowidx(), L_.values(), U_.colptr(),
                         U_.rowidx(), U_.values(), nullptr));
    Check(ipxgpu_split_prepare(ref.ctx, n + m > 0 ? &nonbasic_scale[0] : nullptr,
                               rowperm_inv_.data(), static_cast<Int>(free_positions_.size()),
                               free_positions_.data()));

    OperatorRecord& rec = ipxb200::RecordOf(this);
    rec.kind = OperatorKind::kSplit;
    rec.ref = ref;
    rec.model = &model_;
    rec.time_B = &time_B_;
    rec.time_Bt = &time_Bt_;
    rec.time_NNt = &time_NNt_;
    rec.reprime = nullptr;  // needs the basis: Prepare again
    ipxb200::ClaimState(ref.ctx, ipxb200::StateSlot::kSplit, this);
    prepared_ = true;
}

const Int* SplittedNormalMatrix::colperm() const { return colperm_.data(); }

double SplittedNormalMatrix::time_B() const { return time_B_; }
double SplittedNormalMatrix::time_Bt() const { return time_Bt_; }
double SplittedNormalMatrix::time_NNt() const { return time_NNt_; }

void SplittedNormalMatrix::reset_time() {
    time_B_ = 0.0;
    time_Bt_ = 0.0;
    time_NNt_ = 0.0;
}

// Reference src/splitted_normal_matrix.cc:90-117. A single host-visible apply
// cannot be split into its three device phases without extra synchronisation,
// so its wall time is booked on time_NNt_; inside a CR solve the device timers
// fill all three accumulators (conjugate_residuals_gpu.cc).
void SplittedNormalMatrix::_Apply(const Vector& rhs, Vector& lhs, double* rhs_dot_lhs) {
    assert(prepared_);
    Timer timer;
    OperatorRecord& rec = ipxb200::RecordOf(this);
    if (!ipxb200::StillCurrent(rec))
        throw std::logic_error("SplittedNormalMatrix: device context was rebuilt; call Prepare");
    ipxb200::EnsurePrimed(rec, this);
    if (rhs.size() > 0)
        Check(ipxgpu_split_apply(rec.ref.ctx, &rhs[0], &lhs[0], rhs_dot_lhs));
    else if (rhs_dot_lhs)
        *rhs_dot_lhs = 0.0;
    time_NNt_ += timer.Elapsed();
}

}  // namespace ipx
