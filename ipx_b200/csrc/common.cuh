// Shared device/host definitions for libipxgpu (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ipxgpu {

constexpr int kBlock = 256;       // threads per CTA for every kernel here
constexpr int kWarps = kBlock / 32;
constexpr int kTileNnz = 2048;    // nonzeros staged per CTA (16 KB of products)
constexpr int kTileSeg = 1024;    // max segments (columns / rows) per tile

// One CTA's share of a compressed sparse structure: segments
// [seg0, seg0+nseg) with entries [p0, p1). A segment longer than kTileNnz is
// cut into chunks (long_id >= 0) whose partial sums are combined, in chunk
// order, by whichever CTA finishes last.
struct Tile {
    int seg0, nseg, p0, p1, long_id, chunk;
};

struct LongInfo {
    const int* first;      // [num_long+1] prefix over chunks
    double* partials;      // [total chunks]
    unsigned* counters;    // [num_long], zero between launches
    double* dots;          // [num_long] fused-reduction share of each long segment
    int num_long;
};

// Scratch for the "last CTA finishes the reduction" pattern.
struct Reduce {
    double* partials;      // [3 * max_grid]: sums, second sums, maxima
    unsigned* ticket;      // zero between launches
};

// Values mirrored into pinned host memory so the host can follow a CR solve
// without synchronising the stream.
struct HostMirror {
    long long iter;
    int done;
    int errflag;
    double resnorm;
    int abort;  // written by the host: asks a persistent solve kernel to stop
    int pad;
};

// Device-resident state of one CR solve. Written only by the finalising CTA of
// a kernel, read by the kernels that follow in stream order.
struct CrState {
    double cdot, pdot, alpha, beta, resnorm, rsdot_prev, tol;
    long long iter, maxiter;
    int done, errflag, precond;
    int applies;  // C.Apply count of the solve (persistent kernel; peer-exchange generations)
    unsigned long long t_last;                      // globaltimer of last stamp
    unsigned long long t_op, t_pre, t_vec;          // accumulated ns
    unsigned long long t_B, t_Bt, t_NNt;
    double* hist;
    long long hist_cap;
    HostMirror* mirror;
};

enum ApplyMode : int {
    kApplyPlain = 0,   // write the dot only
    kApplyCrInit = 1,  // cdot = dot, beta = 0
    kApplyCrIter = 2,  // beta = dot/cdot, cdot = dot, iter++
};

enum TimeSlot : int { kSlotNone = 0, kSlotOp, kSlotPre, kSlotVec, kSlotB, kSlotBt, kSlotNNt };

__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Attributes the time since the previous stamp to `slot` (kernels of one solve
// run back to back on one stream, so the previous stamp is the start).
__device__ __forceinline__ void stamp(CrState* st, int slot) {
    const unsigned long long now = globaltimer();
    const unsigned long long dt = now - st->t_last;
    st->t_last = now;
    switch (slot) {
        case kSlotOp: st->t_op += dt; break;
        case kSlotPre: st->t_pre += dt; break;
        case kSlotVec: st->t_vec += dt; break;
        case kSlotB: st->t_B += dt; break;
        case kSlotBt: st->t_Bt += dt; break;
        case kSlotNNt: st->t_NNt += dt; break;
        default: break;
    }
}

// Deterministic CTA-wide sum; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        v = s_red[0];
#pragma unroll
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v += s_red[w];
    }
    return v;
}

__device__ __forceinline__ double block_max(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        v = s_red[0];
#pragma unroll
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = fmax(v, s_red[w]);
    }
    return v;
}

// Publishes this CTA's partials (two sums and one max) and returns true (to
// all threads) in the CTA that arrives last; that CTA then holds the grid-wide
// results in thread 0. Partials are combined in CTA order, so the result does
// not depend on scheduling.
__device__ __forceinline__ bool grid_reduce(const Reduce& red, double my_sum, double my_sum2,
                                            double my_max, double* s_red, int* s_flag,
                                            double* tot_sum, double* tot_sum2,
                                            double* tot_max, const double* extra = nullptr,
                                            int nextra = 0) {
    // `extra`: further addends of the first sum that were published (before
    // the publisher's own ticket) at fixed slots, e.g. by long-segment CTAs.
    const int nblk = gridDim.x;
    if (threadIdx.x == 0) {
        red.partials[blockIdx.x] = my_sum;
        red.partials[nblk + blockIdx.x] = my_sum2;
        red.partials[2 * nblk + blockIdx.x] = my_max;
        __threadfence();
        const unsigned t = atomicAdd(red.ticket, 1u);
        *s_flag = (t == (unsigned)nblk - 1u);
    }
    __syncthreads();
    if (!*s_flag) return false;
    __threadfence();
    double s = 0.0, s2 = 0.0, mx = 0.0;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        s += __ldcg(red.partials + b);
        s2 += __ldcg(red.partials + nblk + b);
        mx = fmax(mx, __ldcg(red.partials + 2 * nblk + b));
    }
    for (int b = threadIdx.x; b < nextra; b += blockDim.x) s += __ldcg(extra + b);
    s = block_sum(s, s_red);
    s2 = block_sum(s2, s_red);
    mx = block_max(mx, s_red);
    if (threadIdx.x == 0) {
        *tot_sum = s;
        *tot_sum2 = s2;
        *tot_max = mx;
        *red.ticket = 0u;
    }
    return true;
}

__device__ __forceinline__ void publish(CrState* st) {
    HostMirror* mir = st->mirror;
    if (!mir) return;
    volatile HostMirror* v = mir;
    v->resnorm = st->resnorm;
    v->errflag = st->errflag;
    v->iter = st->iter;
    __threadfence_system();
    v->done = st->done;
    __threadfence_system();
}

// Scalar step that follows C.Apply in both CR variants
// (reference src/conjugate_residuals.cc:75-81, :176-184).
__device__ __forceinline__ void after_apply(CrState* st, int mode, double dot, int slot) {
    if (mode == kApplyCrInit) {
        st->cdot = dot;
        st->beta = 0.0;
    } else if (mode == kApplyCrIter) {
        st->beta = dot / st->cdot;
        st->cdot = dot;
        st->iter += 1;
    }
    stamp(st, slot);
}

}  // namespace ipxgpu
