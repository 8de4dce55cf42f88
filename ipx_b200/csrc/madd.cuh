// lhs += alpha * AI * rhs  and  lhs += alpha * AI' * rhs over the resident matrix, every sum taken
// in the order of the reference's loops (src/sparse_matrix.cc:194-209 with DotColumn /
// ScatterColumn of src/sparse_matrix.h:136-152), products and additions rounded separately, so the
// result is bit-identical to the host code it replaces: the residuals b - AI*x and c - AI'*y of
// Iterate::ComputeResiduals (src/iterate.cc:543-551) and the starting point's AI'*y
// (src/ipm.cc:191) - two sweeps over AI per interior point iteration.
#pragma once

#include "common.cuh"

namespace ipxgpu {

// 'N': the reference walks the columns j = 0, 1, ... and adds (alpha*rhs[j]) * a_ij to lhs[i], so
// lhs[i] receives its terms in ascending column order starting from its old value - the order of
// the entries of row i in the panel's CSR. One warp per row: the lanes form 32 products at a
// time (coalesced loads, independent), then every lane adds them to the running sum one after
// the other (shuffles; all lanes carry the same sum). The slack column n+i comes last
// (x_slack != nullptr on the last panel): its entry is 1.
__global__ void __launch_bounds__(kBlock)
madd_rows_kernel(int m, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                 const double* __restrict__ val, const double* __restrict__ x, double alpha,
                 double* lhs, const double* __restrict__ x_slack) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kBlock) >> 5;
    for (int i = warp; i < m; i += nwarps) {
        double acc = lhs[i];
        const int p1 = rowptr[i + 1];
        for (int p0 = rowptr[i]; p0 < p1; p0 += 32) {
            const int p = p0 + lane;
            double prod = 0.0;
            if (p < p1) prod = __dmul_rn(__dmul_rn(alpha, __ldcg(x + __ldcs(colidx + p))), __ldcs(val + p));
            const int cnt = min(32, p1 - p0);
            if (cnt == 32) {
#pragma unroll
                for (int l = 0; l < 32; l++)
                    acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, prod, l));
            } else {
                for (int l = 0; l < cnt; l++)
                    acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, prod, l));
            }
        }
        if (x_slack) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(alpha, x_slack[i]), 1.0));
        if (lane == 0) lhs[i] = acc;
    }
}

// 'T', structural columns: lhs[j] += alpha * d with d = sum over the column's entries, in order,
// of rhs[row] * a, starting from 0. One thread per column (a column of a sparse LP holds a
// handful of entries; neighbouring threads read neighbouring entries).
__global__ void __launch_bounds__(kBlock)
madd_cols_kernel(int ncols, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                 const double* __restrict__ val, const double* __restrict__ y, double alpha,
                 double* lhs) {
    for (int j = blockIdx.x * kBlock + threadIdx.x; j < ncols; j += gridDim.x * kBlock) {
        double d = 0.0;
        const int p1 = colptr[j + 1];
        for (int p = colptr[j]; p < p1; p++)
            d = __dadd_rn(d, __dmul_rn(__ldcg(y + __ldcs(rowidx + p)), __ldcs(val + p)));
        lhs[j] = __dadd_rn(lhs[j], __dmul_rn(alpha, d));
    }
}

// 'T', slack columns: column n+i is the unit vector e_i, d = 0 + rhs[i] * 1.
__global__ void __launch_bounds__(kBlock)
madd_slack_cols_kernel(int m, const double* __restrict__ y, double alpha, double* lhs_slack) {
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < m; i += gridDim.x * kBlock) {
        const double d = __dadd_rn(0.0, __dmul_rn(y[i], 1.0));
        lhs_slack[i] = __dadd_rn(lhs_slack[i], __dmul_rn(alpha, d));
    }
}

}  // namespace ipxgpu
