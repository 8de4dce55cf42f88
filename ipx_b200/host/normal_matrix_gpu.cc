// GPU drop-in for ipx::NormalMatrix: defines the members declared in the
// UNMODIFIED reference header src/normal_matrix.h:18-41 and is linked instead
// of src/normal_matrix.cc. Computation runs in libipxgpu (include/ipxgpu.h);
// there is no host fallback.

#include "normal_matrix.h"

#include <cassert>

#include "gpu_bridge.h"
#include "timer.h"

namespace ipx {

using ipxb200::Check;
using ipxb200::OperatorKind;
using ipxb200::OperatorRecord;

NormalMatrix::NormalMatrix(const Model& model) : model_(model) {
    ipxb200::Forget(this);
}

// Reference src/normal_matrix.cc:32-35 captures the pointer only. The weights
// are uploaded here (their only caller rebuilds W right before calling
// Prepare, src/kkt_solver_diag.cc:59); the pointer is kept so that a rebuilt
// device context can be re-primed.
void NormalMatrix::Prepare(const double* W) {
    W_ = W;
    prepared_ = false;
    const ipxb200::ContextRef ref = ipxb200::ContextFor(model_);
    if (!ipxb200::ConsumeWeightsHint(ref.ctx, W)) Check(ipxgpu_normal_prepare(ref.ctx, W));
    OperatorRecord& rec = ipxb200::RecordOf(this);
    rec.kind = OperatorKind::kNormal;
    rec.ref = ref;
    rec.model = &model_;
    rec.time = &time_;
    rec.reprime = [this] { Prepare(W_); };
    ipxb200::ClaimState(ref.ctx, ipxb200::StateSlot::kWeights, this);
    prepared_ = true;
}

double NormalMatrix::time() const { return time_; }

void NormalMatrix::reset_time() { time_ = 0.0; }

// Reference src/normal_matrix.cc:45-126.
void NormalMatrix::_Apply(const Vector& rhs, Vector& lhs, double* rhs_dot_lhs) {
    const Int m = model_.rows();
    Timer timer;
    assert(prepared_);
    assert((Int)lhs.size() == m);
    assert((Int)rhs.size() == m);
    OperatorRecord& rec = ipxb200::RecordOf(this);
    // context rebuilt, or its weights replaced by another instance on this model: prime again
    if (!ipxb200::StillCurrent(rec) ||
        !ipxb200::OwnsState(rec.ref.ctx, ipxb200::StateSlot::kWeights, this))
        Prepare(W_);
    if (m > 0)
        Check(ipxgpu_normal_apply(rec.ref.ctx, &rhs[0], &lhs[0], rhs_dot_lhs));
    else if (rhs_dot_lhs)
        *rhs_dot_lhs = 0.0;
    time_ += timer.Elapsed();
}

}  // namespace ipx
