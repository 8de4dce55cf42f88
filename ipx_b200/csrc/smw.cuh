// Dense-column part of the diagonal preconditioner (reference
// src/diagonal_precond.cc:48-102 Factorize, :133-149 _Apply): with the nd dense columns Ad
// of AI (at most 1000, src/model.cc:34-56) left out of the diagonal E,
//
//     inv(P) v = inv(E) (v - Ad inv(S) Ad' inv(E) v),   S = inv(Wd) + Ad' inv(E) Ad,
//
// where S = L L' has been factorized once per IPM iteration (host LAPACK, as in the
// reference) and L lives in device memory. One preconditioner apply is three kernels that
// stay on the solve's stream, so the CR loop never returns to the host:
//
//   smw_gather   b[k] = sum_i Ad[i,k] * (v[i] / E[i])      column chunks, fixed-order sums
//   smw_solve    z = inv(L L') b                           one CTA, blocked substitutions
//   finish       lhs[i] = (v[i] - sum_k Ad[i,k] z[k]) / E[i] fused with the dot / the CR
//                direction stage's reductions and scalar tests (cr_kernels.cuh)
//
// Up to two right-hand sides ride through one pass (P*Cstep and, every fifth iteration,
// P*residual: reference src/conjugate_residuals.cc:163-167, :187-207).
#pragma once

#include "common.cuh"

namespace ipxgpu {

constexpr int kSmwMaxCols = 1024;   // dense columns (reference: <= 1000)
constexpr int kSmwChunk = 4096;     // entries of a dense column per CTA of the gather
constexpr int kSmwSolveThreads = 1024;

struct SmwDev {
    int nd = 0, m = 0;
    // Ad by column (entries of column k in row order) and by row
    int* colptr = nullptr;
    int* rowidx = nullptr;
    double* colval = nullptr;
    int* rowptr = nullptr;
    int* colidx = nullptr;
    double* rowval = nullptr;
    // gather chunks: chunk c covers entries [chunk_p0[c], chunk_p0[c+1]) of column chunk_col[c]
    int nchunks = 0;
    int* chunk_col = nullptr;
    int* chunk_p0 = nullptr;     // [nchunks + 1]
    int* col_chunk0 = nullptr;   // [nd + 1] first chunk of every column
    double* partials = nullptr;  // [2][nchunks]
    double* L = nullptr;         // nd x nd, column-major, lower triangle
    double* z = nullptr;         // [2][kSmwMaxCols]
    int warp_rows = 0;           // finish: one warp per row (long rows) instead of one thread
};

// Which right-hand sides a pass carries.
enum SmwWhich : int {
    kSmwFirst = 0,       // v0 only
    kSmwRecompute = 1,   // v0, and v1 when the CR state says this is a recompute iteration
    kSmwBoth = 2,        // v0 and v1
};

__device__ __forceinline__ bool smw_second(int which, const CrState* st) {
    if (which == kSmwBoth) return true;
    if (which != kSmwRecompute) return false;
    const long long iter = st->iter;
    return iter > 0 && iter % 5 == 0;
}

// b[k] partials. grid = nchunks.
__global__ void __launch_bounds__(kBlock)
smw_gather_kernel(SmwDev S, const double* __restrict__ E, const double* __restrict__ v0,
                  const double* __restrict__ v1, int which, const CrState* st) {
    __shared__ double s_red[kWarps];
    if (st != nullptr && st->done) return;
    const bool second = st != nullptr ? smw_second(which, st) : which == kSmwBoth;
    const int c = blockIdx.x;
    const int p0 = S.chunk_p0[c], p1 = S.chunk_p0[c + 1];
    double a0 = 0.0, a1 = 0.0;
    for (int p = p0 + threadIdx.x; p < p1; p += kBlock) {
        const int i = S.rowidx[p];
        const double a = S.colval[p];
        const double e = E[i];
        a0 += __dmul_rn(a, v0[i] / e);
        if (second) a1 += __dmul_rn(a, v1[i] / e);
    }
    const double t0 = block_sum(a0, s_red);
    const double t1 = block_sum(a1, s_red);
    if (threadIdx.x == 0) {
        S.partials[c] = t0;
        S.partials[S.nchunks + c] = t1;
    }
}

// Solves L x = b (forward) and L' z = x (backward) in shared memory, 32 columns at a time:
// warp 0 solves the 32 x 32 diagonal block in registers, all threads then fold the block's
// solution into the rest (forward) / all warps first take the solved rest out of the block's
// right-hand sides (backward). x: the CTA's shared vector of nd entries.
__device__ __forceinline__ void smw_chol_solve(const double* __restrict__ L, int nd, double* x,
                                               double* s_t) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x;
    const size_t ld = (size_t)nd;
    // ---- forward: rows of a block ascending, columns ascending per row ----
    for (int kb = 0; kb < nd; kb += 32) {
        const int nb = min(32, nd - kb);
        if (warp == 0) {
            double lrow[32];
#pragma unroll
            for (int k = 0; k < 32; k++)
                lrow[k] = (lane < nb && k <= lane) ? L[(size_t)(kb + lane) + (size_t)(kb + k) * ld] : 1.0;
            double b = lane < nb ? x[kb + lane] : 0.0;
#pragma unroll
            for (int k = 0; k < 32; k++) {
                if (lane == k) b = b / lrow[k];
                const double xk = __shfl_sync(0xffffffffu, b, k);
                if (lane > k) b = b - __dmul_rn(lrow[k], xk);
            }
            if (lane < nb) x[kb + lane] = b;
        }
        __syncthreads();
        for (int i = kb + 32 + tid; i < nd; i += nthr) {
            double acc = x[i];
#pragma unroll 8
            for (int k = 0; k < 32; k++)
                acc = acc - __dmul_rn(L[(size_t)i + (size_t)(kb + k) * ld], x[kb + k]);
            x[i] = acc;
        }
        __syncthreads();
    }
    // ---- backward ----
    const int last = ((nd - 1) / 32) * 32;
    for (int kb = last; kb >= 0; kb -= 32) {
        const int nb = min(32, nd - kb);
        // warp j: t[j] = sum over the solved rows below the block of L[i, kb+j] * z[i]
        for (int j = warp; j < nb; j += nthr >> 5) {
            double acc = 0.0;
            const double* col = L + (size_t)(kb + j) * ld;
            for (int i = kb + 32 + lane; i < nd; i += 32) acc += __dmul_rn(col[i], x[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
            if (lane == 0) s_t[j] = acc;
        }
        __syncthreads();
        if (warp == 0) {
            double lcol[32];  // lcol[k] = L[kb+k, kb+lane], k >= lane
#pragma unroll
            for (int k = 0; k < 32; k++)
                lcol[k] = (k < nb && k >= lane) ? L[(size_t)(kb + k) + (size_t)(kb + lane) * ld] : 1.0;
            double b = lane < nb ? x[kb + lane] - s_t[lane] : 0.0;
#pragma unroll
            for (int k = 31; k >= 0; k--) {
                if (lane == k) b = b / lcol[k];
                const double zk = __shfl_sync(0xffffffffu, b, k);
                if (lane < k && k < nb) b = b - __dmul_rn(lcol[k], zk);
            }
            if (lane < nb) x[kb + lane] = b;
        }
        __syncthreads();
    }
}

// One CTA: b = sum of the chunk partials in chunk order, z = inv(L L') b for one or two
// right-hand sides.
__global__ void __launch_bounds__(kSmwSolveThreads)
smw_solve_kernel(SmwDev S, int which, const CrState* st) {
    __shared__ double s_x[kSmwMaxCols];
    __shared__ double s_t[32];
    if (st != nullptr && st->done) return;
    const bool second = st != nullptr ? smw_second(which, st) : which == kSmwBoth;
    for (int r = 0; r < (second ? 2 : 1); r++) {
        for (int k = threadIdx.x; k < S.nd; k += blockDim.x) {
            double b = 0.0;
            for (int c = S.col_chunk0[k]; c < S.col_chunk0[k + 1]; c++)
                b += S.partials[(size_t)r * S.nchunks + c];
            s_x[k] = b;
        }
        __syncthreads();
        smw_chol_solve(S.L, S.nd, s_x, s_t);
        for (int k = threadIdx.x; k < S.nd; k += blockDim.x) S.z[(size_t)r * kSmwMaxCols + k] = s_x[k];
        __syncthreads();
    }
}

// (v[i] - sum_k Ad[i,k] z[k]) / E[i] for the rows this thread (or warp) owns; calls
// fn(i, value0, value1). The row sum runs in column order (thread per row: the reference's
// DotColumn order, src/diagonal_precond.cc:144-146).
template <class F>
__device__ __forceinline__ void smw_rows(const SmwDev& S, const double* __restrict__ E,
                                         const double* __restrict__ v0,
                                         const double* __restrict__ v1, bool second, F fn) {
    const double* z0 = S.z;
    const double* z1 = S.z + kSmwMaxCols;
    if (!S.warp_rows) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.m; i += gridDim.x * blockDim.x) {
            double c0 = 0.0, c1 = 0.0;
            for (int q = S.rowptr[i]; q < S.rowptr[i + 1]; q++) {
                const int k = S.colidx[q];
                const double a = S.rowval[q];
                c0 += __dmul_rn(a, z0[k]);
                if (second) c1 += __dmul_rn(a, z1[k]);
            }
            const double e = E[i];
            fn(i, (v0[i] - c0) / e, second ? (v1[i] - c1) / e : 0.0);
        }
    } else {
        const int lane = threadIdx.x & 31;
        const int wpg = (gridDim.x * blockDim.x) >> 5;
        for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < S.m; i += wpg) {
            double c0 = 0.0, c1 = 0.0;
            for (int q = S.rowptr[i] + lane; q < S.rowptr[i + 1]; q += 32) {
                const int k = S.colidx[q];
                const double a = S.rowval[q];
                c0 += __dmul_rn(a, z0[k]);
                if (second) c1 += __dmul_rn(a, z1[k]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c0 += __shfl_down_sync(0xffffffffu, c0, o);
                c1 += __shfl_down_sync(0xffffffffu, c1, o);
            }
            if (lane == 0) {
                const double e = E[i];
                fn(i, (v0[i] - c0) / e, second ? (v1[i] - c1) / e : 0.0);
            }
        }
    }
}

// lhs = inv(P) rhs with the fused dot (stand-alone preconditioner apply).
__global__ void __launch_bounds__(kBlock)
smw_apply_finish_kernel(SmwDev S, const double* __restrict__ E, const double* __restrict__ rhs,
                        double* __restrict__ lhs, Reduce red, double* dot_out) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    double acc = 0.0;
    smw_rows(S, E, rhs, nullptr, false, [&](int i, double l, double) {
        lhs[i] = l;
        acc += __dmul_rn(l, rhs[i]);
    });
    const double b = block_sum(acc, s_red);
    double ts, ts2, tm;
    if (grid_reduce(red, b, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) && threadIdx.x == 0)
        *dot_out = ts;
}

}  // namespace ipxgpu
