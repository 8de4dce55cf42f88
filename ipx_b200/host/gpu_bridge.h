// Side tables that attach device state to IPX objects whose class layouts are
// fixed by the unmodified reference headers.
//
//  * one ipxgpu context per ipx::Model, created lazily, validated by
//    (AI.values(), AI.entries(), rows, cols) plus a content fingerprint, and
//    kept in a small LRU cache (IPXGPU_MAX_CONTEXTS, default 4) because Model
//    has no destructor hook;
//  * one record per LinearOperator instance (NormalMatrix, DiagonalPrecond,
//    SplittedNormalMatrix) telling ConjugateResiduals which device operator it
//    is and where its time accumulators live.

#ifndef IPXB200_GPU_BRIDGE_H_
#define IPXB200_GPU_BRIDGE_H_

#include <functional>

#include "ipxgpu.h"
#include "linear_operator.h"
#include "model.h"

namespace ipxb200 {

struct ContextRef {
    ipxgpu_ctx* ctx;
    unsigned long long generation;  // changes when the context is rebuilt
    ContextRef() : ctx(nullptr), generation(0) {}
    ContextRef(ipxgpu_ctx* c, unsigned long long g) : ctx(c), generation(g) {}
};

// Validated lookup; creates (and uploads AI) on first use or when the model's
// matrix changed. Throws std::bad_alloc or std::runtime_error.
ContextRef ContextFor(const ipx::Model& model);
// Cheap lookup by address only; {nullptr, 0} if there is no context.
ContextRef CurrentContext(const ipx::Model& model);

// The single-GPU context whose model's AI is exactly this matrix object's data (same arrays, same
// dimensions, same content fingerprint), or nullptr: ipx::MultiplyAdd's way from a SparseMatrix
// back to the device copy (multiply_add_gpu.cc). Never creates a context.
ipxgpu_ctx* ContextOfMatrix(const ipx::SparseMatrix& A);

// Several GPUs behind LpSolver: with IPXGPU_NGPUS = G > 1 in the environment KKTSolverDiag
// runs on a group of G column-sharded contexts of this process (ipxgpu_create_group: one host
// thread per GPU for the duration of a call, NVLink peer exchange inside the CR kernels), the
// diagonal-preconditioned phase being where the column-sharded A*D^2*A' of SURVEY.md
// section 8e applies. The basis-preconditioned phase keeps the single-GPU context
// (its triangular solves are sequential, section 8e "not sharded").
int GroupSize();  // G, or 1
ContextRef GroupContextFor(const ipx::Model& model);
ContextRef CurrentGroupContext(const ipx::Model& model);

// Maps a C-ABI return code to IPX's exception convention
// (reference src/lp_solver.cc:98-105): out of memory -> std::bad_alloc,
// anything else -> std::runtime_error carrying ipxgpu_last_error().
void Check(int rc);

enum class OperatorKind { kNone, kNormal, kDiagonal, kSplit };

struct OperatorRecord {
    OperatorKind kind = OperatorKind::kNone;
    ContextRef ref;
    const ipx::Model* model = nullptr;
    double* time = nullptr;     // NormalMatrix::time_ / DiagonalPrecond::time_
    double* time_B = nullptr;   // SplittedNormalMatrix accumulators
    double* time_Bt = nullptr;
    double* time_NNt = nullptr;
    // Puts this operator's state back on the device when another instance on the same model
    // has primed the context since (empty: the operator cannot; its user must prepare again).
    std::function<void()> reprime;
};

// A context holds ONE set of weights, ONE diagonal, ONE set of factors ... while the reference's
// classes each own theirs. Whoever primes a part of the context claims it; before an object
// uses the device state it checks that the claim is still its own and re-primes (or refuses)
// otherwise, so two live instances on one Model cannot silently use each other's state.
enum class StateSlot : int { kWeights = 0, kDiagonal, kSplit, kKktDiag, kCount };
void ClaimState(ipxgpu_ctx* ctx, StateSlot slot, const void* owner);
bool OwnsState(ipxgpu_ctx* ctx, StateSlot slot, const void* owner);
// Re-primes the operator behind `rec` if its claim was taken; throws std::logic_error when
// the operator cannot re-prime itself.
void EnsurePrimed(OperatorRecord& rec, const ipx::LinearOperator* op);

// Record of an operator instance (keyed by its LinearOperator base address).
OperatorRecord& RecordOf(const ipx::LinearOperator* op);
// Lookup without creating; nullptr if unknown.
OperatorRecord* FindRecord(const ipx::LinearOperator* op);
// Drops a stale record (called by the constructors of the drop-in classes).
void Forget(const ipx::LinearOperator* op);
// True if the record's context is still the model's current context.
bool StillCurrent(const OperatorRecord& rec);

// One-shot hints set by KKTSolverDiag::_Factorize when weights and diagonal
// were built on the device, so that the member operators' Prepare/Factorize do
// not upload/rebuild them again.
void SetResidentHint(ipxgpu_ctx* ctx, const double* W);
bool ConsumeWeightsHint(ipxgpu_ctx* ctx, const double* W);
bool ConsumeDiagonalHint(ipxgpu_ctx* ctx, const double* W);

}  // namespace ipxb200

#endif  // IPXB200_GPU_BRIDGE_H_
