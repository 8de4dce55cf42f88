// ipx::MultiplyAdd (reference src/sparse_matrix.h:154-157, src/sparse_matrix.cc:194-209) for the
// one matrix that is resident on the device: a model's AI. The interior point method calls it
// twice per iteration for the residuals b - AI*x and c - AI'*y (src/iterate.cc:543-551) and once
// for the starting point (src/ipm.cc:191); with the KKT solve on the device these two sweeps
// over AI on one host core are the largest single item left of an iteration (SURVEY.md
// section 8f-1, "residual SpMVs").
//
// The seam: src/sparse_matrix.cc is compiled UNCHANGED but with -DMultiplyAdd=MultiplyAdd_reference
// (ipx_b200/build.py), which renames this one free function in that translation unit; every
// other translation unit calls ipx::MultiplyAdd as before and binds to the definition below.
// For a matrix that is some model's AI with a live single-GPU context the product runs in
// libipxgpu (ipxgpu_multiply_add: sums in the reference's order, products and additions rounded
// separately - bit-identical); every other matrix (a basis matrix in the LU residual checks, the
// user's unscaled A) is not on the device and goes to the reference's own function.

#include "sparse_matrix.h"

#include <cassert>

#include "gpu_bridge.h"

namespace ipx {

// the reference's definition under its compile-time name (ref_sparse_matrix_renamed.o)
void MultiplyAdd_reference(const SparseMatrix& A, const Vector& rhs, double alpha, Vector& lhs,
                           char trans);

void MultiplyAdd(const SparseMatrix& A, const Vector& rhs, double alpha, Vector& lhs, char trans) {
    // an empty side (no rows, or no columns at all) leaves nothing to do on the device
    ipxgpu_ctx* ctx = rhs.size() == 0 || lhs.size() == 0 ? nullptr : ipxb200::ContextOfMatrix(A);
    if (!ctx) {
        MultiplyAdd_reference(A, rhs, alpha, lhs, trans);
        return;
    }
    const Int m = A.rows();
    const Int n = A.cols();
    if (trans == 't' || trans == 'T') {
        assert((Int)rhs.size() == m);
        assert((Int)lhs.size() == n);
    } else {
        assert((Int)rhs.size() == n);
        assert((Int)lhs.size() == m);
    }
    ipxb200::Check(ipxgpu_multiply_add(ctx, &rhs[0], alpha, &lhs[0], trans));
}

}  // namespace ipx
