// Column sweeps of the maximum-volume heuristic (reference src/maxvolume.cc:202-320).
//
// Maxvolume::Driver keeps one weight per column of AI = [A I] and, per basis update, (a) picks
// the two largest weights (FindLargest, :170-200), (b) forms the tableau row of the leaving
// variable, row[j] = AI[:,j]' * btran over the nonbasic columns (Basis::TableauRow, dense branch,
// src/basis.cc:266-279) and (c) adds alpha * row[j] * colscale[j] to every weight (:302-307).
// All three are sweeps over the n+m columns; the pivoting around them (two solves with the
// factorization and the stability test per update) stays with the host's Basis object.
// colscale and colweights live on the device for the duration of a run; per update the host sends
// btran (m doubles) and receives the two candidates.
//
// The structural columns go through the generic segmented sweep (spmv_kernels.cuh): columns of
// up to 16 entries are summed by one lane in storage order, i.e. in the order of the
// reference's DotColumn. The slack columns and the search for the two largest weights are one
// streaming kernel over all n+m weights.
#pragma once

#include "common.cuh"

namespace ipxgpu {

// Initial weights (:220-231): colweights[j] = (AI[:,j]'work) * colscale[j] where colscale[j] != 0.
struct OpColMaxvolInit {
    static constexpr bool kReduce = false;
    const double* work;
    const double* colscale;
    double* colweights;
    __device__ __forceinline__ double prod(int i, double a) const {
        return __dmul_rn(a, IPXGPU_GATHER(work + i));
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        const double cs = colscale[seg];
        colweights[seg] = cs != 0.0 ? __dmul_rn(sum, cs) : 0.0;
        return 0.0;
    }
    __device__ __forceinline__ void finalize(double, CrState*) const {}
};

// Weight update (:302-305): colweights[j] += alpha * row[j] * colscale[j]. Columns with
// colscale[j] == 0 (basic, fixed or skipped ones) keep their weight, as in the reference where
// the product is an exact zero.
struct OpColMaxvolUpdate {
    static constexpr bool kReduce = false;
    const double* btran;
    const double* colscale;
    double* colweights;
    double alpha;
    __device__ __forceinline__ double prod(int i, double a) const {
        return __dmul_rn(a, IPXGPU_GATHER(btran + i));
    }
    __device__ __forceinline__ double epilogue(int seg, double sum) const {
        const double cs = colscale[seg];
        if (cs != 0.0)
            colweights[seg] = __dadd_rn(colweights[seg], __dmul_rn(__dmul_rn(alpha, sum), cs));
        return 0.0;
    }
    __device__ __forceinline__ void finalize(double, CrState*) const {}
};

// The two best (|weight|, column) pairs of a set: larger |weight| first, the smaller column
// among equal ones - the pair FindLargest's ascending scan with strict comparisons ends with.
// Only positive weights count; col < 0: empty.
struct Top2 {
    double w1, w2;
    long long j1, j2;
};

__device__ __forceinline__ bool mv_better(double wa, long long ja, double wb, long long jb) {
    if (ja < 0) return false;
    if (jb < 0) return true;
    return wa > wb || (wa == wb && ja < jb);
}

__device__ __forceinline__ void mv_push(Top2& t, double w, long long j) {
    if (!(w > 0.0)) return;  // zeros and NaNs never become candidates (:181-197)
    if (mv_better(w, j, t.w1, t.j1)) {
        t.w2 = t.w1;
        t.j2 = t.j1;
        t.w1 = w;
        t.j1 = j;
    } else if (mv_better(w, j, t.w2, t.j2)) {
        t.w2 = w;
        t.j2 = j;
    }
}

__device__ __forceinline__ void mv_merge(Top2& t, const Top2& o) {
    if (o.j1 >= 0) mv_push(t, o.w1, o.j1);
    if (o.j2 >= 0) mv_push(t, o.w2, o.j2);
}

struct MaxvolArgs {
    long long n, m;
    const double* vec;   // work (init) or btran (update), m entries; nullptr: search only
    double* colscale;    // n+m
    double* colweights;  // n+m
    int update;          // 0: initial weights of the slack columns, 1: update them
    double alpha;
    long long jb, jn;    // update: entering / leaving the nonbasic set (-1: none)
    double cw_jb;        // update: weight the reference assigns to jb (:306)
    Top2* partials;      // [gridDim.x]
    unsigned* ticket;    // zero between launches
    Top2* out;
};

// Slack columns (AI[:,n+i] = e_i, so the dot is vec[i]) + the two assignments that close an
// update (colweights[jb], colweights[jn], :306-307) + FindLargest over all n+m weights.
__global__ void __launch_bounds__(kBlock) maxvol_finish_kernel(MaxvolArgs a) {
    __shared__ Top2 s_top[kWarps];
    __shared__ int s_last;
    Top2 best{0.0, 0.0, -1, -1};
    const long long total = a.n + a.m;
    const long long stride = (long long)gridDim.x * kBlock;
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < total; j += stride) {
        double w;
        if (j >= a.n && a.vec != nullptr) {
            const double cs = a.colscale[j];
            const double dot = a.vec[j - a.n];
            if (!a.update) {
                w = cs != 0.0 ? __dmul_rn(dot, cs) : 0.0;
            } else {
                w = a.colweights[j];
                if (cs != 0.0) w = __dadd_rn(w, __dmul_rn(__dmul_rn(a.alpha, dot), cs));
            }
            if (a.update && j == a.jb) w = a.cw_jb;
            if (a.update && j == a.jn) w = 0.0;
            a.colweights[j] = w;
        } else {
            w = __ldcs(a.colweights + j);
            if (a.update && a.vec != nullptr && (j == a.jb || j == a.jn)) {
                w = j == a.jb ? a.cw_jb : 0.0;
                a.colweights[j] = w;
            }
        }
        mv_push(best, fabs(w), j);
    }
    // warp, then CTA, then grid: the order of merging does not matter ((|w|, j) is a total order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Top2 other;
        other.w1 = __shfl_down_sync(0xffffffffu, best.w1, o);
        other.w2 = __shfl_down_sync(0xffffffffu, best.w2, o);
        other.j1 = __shfl_down_sync(0xffffffffu, best.j1, o);
        other.j2 = __shfl_down_sync(0xffffffffu, best.j2, o);
        mv_merge(best, other);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_top[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kWarps; w++) mv_merge(best, s_top[w]);
        a.partials[blockIdx.x] = best;
        __threadfence();
        const unsigned t = atomicAdd(a.ticket, 1u);
        s_last = (t == gridDim.x - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    Top2 all{0.0, 0.0, -1, -1};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kBlock) {
        Top2 p;
        p.w1 = __ldcg(&a.partials[b].w1);
        p.w2 = __ldcg(&a.partials[b].w2);
        p.j1 = __ldcg(&a.partials[b].j1);
        p.j2 = __ldcg(&a.partials[b].j2);
        mv_merge(all, p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Top2 other;
        other.w1 = __shfl_down_sync(0xffffffffu, all.w1, o);
        other.w2 = __shfl_down_sync(0xffffffffu, all.w2, o);
        other.j1 = __shfl_down_sync(0xffffffffu, all.j1, o);
        other.j2 = __shfl_down_sync(0xffffffffu, all.j2, o);
        mv_merge(all, other);
    }
    __syncthreads();
    if (lane == 0) s_top[warp] = all;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kWarps; w++) mv_merge(all, s_top[w]);
        *a.out = all;
        *a.ticket = 0u;
    }
}

// colscale[j] = value for up to four columns (the exchange of an update, a skipped column).
struct MaxvolPoke {
    long long j[4];
    double scale[4];
    int zero_weight[4];
    int count;
};

__global__ void maxvol_poke_kernel(MaxvolPoke p, double* colscale, double* colweights) {
    const int k = threadIdx.x;
    if (k < p.count && p.j[k] >= 0) {
        colscale[p.j[k]] = p.scale[k];
        if (p.zero_weight[k]) colweights[p.j[k]] = 0.0;
    }
}

}  // namespace ipxgpu
