// Segmented gather-reduce sweeps over a compressed sparse structure.
//
// One kernel template serves every sweep of the path; an Op supplies the
// per-entry product and the per-segment epilogue:
//   sweep 1  t = W .* (A' x)          CSC, segment = column  (normal_matrix.cc:67-73)
//   sweep 2  y = Ws.*x + A t, x'y     CSR, segment = row     (normal_matrix.cc:65-66,74-75,123)
//   diag     d = Ws + sum W a^2       CSR                    (diagonal_precond.cc:28-37)
//   rhs / recovery sweeps of KKTSolverDiag                   (kkt_solver_diag.cc:90-92,108-117)
//
// A CTA streams its tile's indices and values with coalesced, evict-first
// loads, gathers the vector through the read-only path, stages the products in
// shared memory and then reduces each segment with 1..32 lanes (chosen from the
// tile's segment count) in a fixed order, so results are run-to-run
// deterministic. The intermediate vector t stays L2-resident between sweep 1
// and sweep 2 of a column panel.
#pragma once

#include "common.cuh"

// Random 8-byte gathers of the generic sweeps: __ldcg (L2 only) instead of __ldg (read-only
// path, L1 allocating): -9 % on config 5 (1M x 20M: 1852 -> 1686 us per apply), where an L1 line
// is never reused before it is evicted.
#ifndef IPXGPU_GATHER
#define IPXGPU_GATHER __ldcg
#endif

namespace ipxgpu {

template <class Op>
__global__ void __launch_bounds__(kBlock)
seg_sweep_kernel(Op op, const Tile* __restrict__ tiles, const int* __restrict__ ptr,
                 const int* __restrict__ idx, const double* __restrict__ val, LongInfo li,
                 Reduce red, CrState* st) {
    __shared__ double s_prod[kTileNnz];
    __shared__ int s_ptr[kTileSeg + 1];
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;

    if (st != nullptr && st->done) return;

    const Tile tile = tiles[blockIdx.x];
    const int tid = threadIdx.x;
    const int nnzT = tile.p1 - tile.p0;
    const int* tidx = idx + tile.p0;
    const double* tval = val + tile.p0;
    double acc = 0.0;  // this thread's share of the fused scalar reduction

    if (tile.nseg == 1) {
        // One (chunk of a) segment for the whole CTA: accumulate in registers.
        double sum = 0.0;
#pragma unroll 4
        for (int k = tid; k < nnzT; k += kBlock)
            sum += op.prod(__ldcs(tidx + k), __ldcs(tval + k));
        sum = block_sum(sum, s_red);
        if (tid == 0) {
            if (tile.long_id < 0) {
                acc += op.epilogue(tile.seg0, sum);
            } else {
                const int base = li.first[tile.long_id];
                const int nch = li.first[tile.long_id + 1] - base;
                li.partials[base + tile.chunk] = sum;
                __threadfence();
                const unsigned t = atomicAdd(li.counters + tile.long_id, 1u);
                if (t == (unsigned)nch - 1u) {
                    __threadfence();
                    double tot = 0.0;
                    for (int c = 0; c < nch; c++) tot += __ldcg(li.partials + base + c);
                    li.counters[tile.long_id] = 0u;
                    // Fixed slot, so the fused sum does not depend on which
                    // chunk's CTA happened to finish last.
                    li.dots[tile.long_id] = op.epilogue(tile.seg0, tot);
                }
            }
        }
    } else {
        for (int s = tid; s <= tile.nseg; s += kBlock) s_ptr[s] = ptr[tile.seg0 + s] - tile.p0;
        constexpr int kPer = kTileNnz / kBlock;
        int ii[kPer];
        double vv[kPer];
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            const int k = tid + u * kBlock;
            if (k < nnzT) {
                ii[u] = __ldcs(tidx + k);
                vv[u] = __ldcs(tval + k);
            }
        }
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            const int k = tid + u * kBlock;
            if (k < nnzT) s_prod[k] = op.prod(ii[u], vv[u]);
        }
        __syncthreads();
        // Lanes per segment: largest power of two with nseg*G <= kBlock.
        int G = 1;
        while (G < 32 && tile.nseg * G * 2 <= kBlock) G <<= 1;
        const int lane = tid & (G - 1);
        const int grp = tid / G;
        const int ngrp = kBlock / G;
        for (int s0 = 0; s0 < tile.nseg; s0 += ngrp) {
            const int s = s0 + grp;
            const bool valid = s < tile.nseg;
            double sum = 0.0;
            if (valid) {
                const int e = s_ptr[s + 1];
                for (int k = s_ptr[s] + lane; k < e; k += G) sum += s_prod[k];
            }
            for (int o = G >> 1; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, G);
            if (valid && lane == 0) acc += op.epilogue(tile.seg0 + s, sum);
        }
    }

    if (Op::kReduce) {
        const double mine = block_sum(acc, s_red);
        double tot_sum, tot_sum2, tot_max;
        if (grid_reduce(red, mine, 0.0, 0.0, s_red, &s_flag, &tot_sum, &tot_sum2, &tot_max,
                        li.dots, li.num_long)) {
            if (tid == 0) op.finalize(tot_sum, st);
        }
    }
}

// ---- Ops ----

// Sweep 1: t[j] = W[j] * sum_p x[row_p] a_p.
struct OpColDotScale {
    static constexpr bool kReduce = false;
    const double* x;
    const double* W;  // this shard's structural weights, or nullptr (W = 1)
    double* t;
    __device__ __forceinline__ double prod(int i, double a) const {
        return __dmul_rn(IPXGPU_GATHER(x + i), a);
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        t[seg] = W ? __dmul_rn(sum, W[seg]) : sum;
        return 0.0;
    }
    __device__ __forceinline__ void finalize(double, CrState*) const {}
};

// Sweep 2: y[i] = (first panel ? slack[i]*x[i] : y[i]) + sum_p t[col_p] a_p,
// with the fused partial dot x'y on the last panel.
struct OpRowGather {
    static constexpr bool kReduce = true;
    const double* t;
    const double* x;
    const double* Ws;  // slack weights W[n..n+m) or nullptr (0); only on rank 0
    double* y;         // m+1 entries; y[m] receives x'y
    int m;
    int first_panel, last_panel;
    int mode;          // ApplyMode, applied when last_panel
    int slot;
    __device__ __forceinline__ double prod(int j, double a) const {
        return __dmul_rn(IPXGPU_GATHER(t + j), a);
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        double v;
        if (first_panel) v = (Ws ? __dmul_rn(x[seg], Ws[seg]) : 0.0) + sum;
        else v = y[seg] + sum;
        y[seg] = v;
        return last_panel ? __dmul_rn(x[seg], v) : 0.0;
    }
    __device__ __forceinline__ void finalize(double total, CrState* st) const {
        if (!last_panel) return;
        y[m] = total;
        if (st) after_apply(st, mode, total, slot);
    }
};

// Spilled segments of a banded sweep (context.cuh Spill). The banded kernel has written its
// result for every segment with an empty sum for these; the generic kernel fills them in.
// Sweep 1: t[map[k]] = W[map[k]] * sum.
struct OpColDotScaleSpill {
    static constexpr bool kReduce = false;
    const double* x;
    const double* W;
    double* t;
    const int* map;
    __device__ __forceinline__ double prod(int i, double a) const {
        return __dmul_rn(IPXGPU_GATHER(x + i), a);
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        const int g = map[seg];
        t[g] = W ? __dmul_rn(sum, W[g]) : sum;
        return 0.0;
    }
    __device__ __forceinline__ void finalize(double, CrState*) const {}
};

// Sweep 2: y[map[k]] += sum; y[m] (= x'y so far) += x[map[k]] * sum, then the scalar step the
// banded kernel left to this one.
struct OpRowGatherSpill {
    static constexpr bool kReduce = true;
    const double* t;
    const double* x;
    double* y;  // m+1 entries
    const int* map;
    int m;
    int mode, slot;
    __device__ __forceinline__ double prod(int j, double a) const {
        return __dmul_rn(IPXGPU_GATHER(t + j), a);
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        const int g = map[seg];
        y[g] = y[g] + sum;
        return __dmul_rn(x[g], sum);
    }
    __device__ __forceinline__ void finalize(double total, CrState* st) const {
        const double dot = y[m] + total;
        y[m] = dot;
        if (st) after_apply(st, mode, dot, slot);
    }
};

// Diagonal build: d[i] = (first ? Ws[i] : d[i]) + sum_p W[col_p] a_p^2.
struct OpRowDiag {
    static constexpr bool kReduce = false;
    const double* W;   // structural weights or nullptr (1)
    const double* Ws;  // slack weights or nullptr (0)
    double* d;
    int first_panel;
    __device__ __forceinline__ double prod(int j, double a) const {
        return W ? __dmul_rn(__dmul_rn(a, IPXGPU_GATHER(W + j)), a) : __dmul_rn(a, a);
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        d[seg] = (first_panel ? (Ws ? Ws[seg] : 0.0) : d[seg]) + sum;
        return 0.0;
    }
    __device__ __forceinline__ void finalize(double, CrState*) const {}
};

// KKT rhs: r[i] = (first ? Ws[i]*as[i] - b[i] : r[i]) + sum_p u[col_p] a_p,
// with u = W.*a on structurals (kkt_solver_diag.cc:90-92).
struct OpRowAffine {
    static constexpr bool kReduce = false;
    const double* u;
    const double* init;  // first-panel initial value per row
    double* r;
    double sign;         // +1: init + sum, -1: init - sum
    int first_panel;
    __device__ __forceinline__ double prod(int j, double a) const {
        return __dmul_rn(IPXGPU_GATHER(u + j), a);
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        const double base = first_panel ? init[seg] : r[seg];
        r[seg] = sign > 0 ? base + sum : base - sum;
        return 0.0;
    }
    __device__ __forceinline__ void finalize(double, CrState*) const {}
};

// KKT recovery: xj = W[j] * (a[j] - A[:,j]'y) (kkt_solver_diag.cc:111-112).
struct OpColRecover {
    static constexpr bool kReduce = false;
    const double* y;
    const double* W;
    const double* a;
    double* xout;
    __device__ __forceinline__ double prod(int i, double v) const {
        return __dmul_rn(IPXGPU_GATHER(y + i), v);
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        xout[seg] = __dmul_rn(W[seg], a[seg] - sum);
        return 0.0;
    }
    __device__ __forceinline__ void finalize(double, CrState*) const {}
};

}  // namespace ipxgpu
