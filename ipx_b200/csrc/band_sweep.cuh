// Banded gather-reduce sweep: out[s] = sum_{e in segment s} v[idx_e] * a_e with
// the gathered vector staged in shared memory.
//
// Why: on B200 a divergent 8-byte global gather costs ~2 L1 wavefront cycles
// per lane, which caps a sweep of A (normal_matrix.cc:67-75) near 20 % of HBM
// bandwidth. A shared-memory gather costs a fraction of that. The gathered
// vector does not fit in shared memory, so the matrix is re-tiled in two
// dimensions at context creation:
//
//   bands  of the gather index space (VB entries, staged by TMA bulk copies
//          into a ring of NBUF shared-memory buffers)
//   blocks of the segment space      (SB accumulators in shared memory)
//
// An item = (segment block, run of consecutive bands) and is one CTA: one
// producer warp that issues the bulk copies (cp.async.bulk, mbarrier
// complete_tx) and NW consumer warps. A segment belongs to ONE consumer warp
// for the whole item (local segment % NW), so every addition to its
// accumulator is issued by that warp in program order and the warps never
// synchronise with each other: a warp waits only for the band it needs
// (full mbarrier) and releases the one it has finished (empty mbarrier).
// Inside a tile (block, band) a segment's entries form one run; a warp's runs
// are dealt to its 32 lanes, longest first to the least loaded lane, and the
// warp's share of the tile is stored as rows of 32 entries (32 keys, 32
// values: 384 contiguous bytes, fully coalesced). The rows of all the item's
// tiles are contiguous per warp, so the warp streams them through two
// register batches of D rows (one in flight while the other is consumed)
// without caring about tile boundaries. A lane sums a run in a register and
// adds it to the segment's accumulator at the run's last entry: no shuffles,
// no atomics, fixed summation order.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace ipxgpu {

constexpr int kBandMaxVB = 32768;        // 15-bit band index
constexpr unsigned kBandLast = 0x8000u;  // key flag: last entry of its run
constexpr int kBandRowBytes = 384;       // 32 keys + 32 values
constexpr int kBandMaxSteps = 256;       // bands per item (row table in shared memory)
constexpr int kBandMaxRun = 64;          // longest run of a segment inside one tile
constexpr int kBandTailRows = 64;        // spare rows after the last one (>= 3*D)

struct BandPlan {
    int V = 0, S = 0;      // gather-vector length, number of segments
    int VB = 0, SB = 0;    // band / block extents
    int NVB = 0, NSB = 0;  // number of bands / blocks
    int K = 0;             // bands per item
    int nparts = 0;        // ceil(NVB / K): partial outputs per segment
    int nitems = 0;        // NSB * nparts
    int NW = 0;            // consumer warps per CTA
    int NBUF = 2;          // band buffers
    long long nnz = 0;
    size_t smem = 0;
};

struct BandDev {
    BandPlan plan;
    int* row_ptr = nullptr;          // [nitems * NW * (K+1)]
    unsigned char* stream = nullptr; // rows * 384 bytes
    double* partials = nullptr;      // [nparts * S] when nparts > 1
    long long rows = 0;
    int debug = 0;
    unsigned long long* trace = nullptr;  // tuning only: [nitems][2*NW + 4] globaltimer stamps
};

struct BandHost {
    std::vector<int> row_ptr;
    std::vector<uint32_t> stream;  // rows * 96 words
    long long rows = 0;
    long long pad_entries = 0;
};

enum BandMode : int {
    kBandColScale = 0,  // out[s] = W ? acc*W[s] : acc
    kBandRowFinal = 1,  // out[s] = (Ws ? x[s]*Ws[s] : 0) + acc, fused dot
    kBandPartial = 2,   // partials[part*S + s] = acc
};

struct BandArgs {
    const double* v;   // gather vector (16-byte aligned)
    const double* W;   // kBandColScale: weights or nullptr
    const double* Ws;  // kBandRowFinal: slack weights or nullptr
    const double* x;   // kBandRowFinal: rhs for slack term and dot
    double* out;       // t (col mode) or y (row mode, m+1 entries)
    int apply_mode;    // ApplyMode for the fused scalar step
    int slot;
};

// Readiness of the gathered vector, piece by piece (persistent CR kernel): piece f = gather
// indices [f * div, (f+1) * div) may be staged once flags[f] has reached gen. The producer lane
// waits for the pieces a band overlaps before it issues the band's bulk copies, so the sweep
// that writes the vector and the sweep that gathers from it need no grid barrier between them.
struct BandReady {
    const unsigned* flags = nullptr;  // nullptr: the whole vector is ready
    unsigned gen = 0;
    int div = 1;
    int nflags = 0;
};

// ---- PTX helpers (mbarrier + bulk copy) ----

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// Bulk copy global -> shared with an L2 eviction policy: the gathered vector is
// re-read by every CTA while the matrix streams through L2, so it is kept with
// evict_last priority.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---- kernel ----
//
// Dynamic shared memory layout (bytes):
//   [0, 64)                      full[NBUF], empty[NBUF] mbarriers (NBUF <= 4)
//   [64, 64 + 4*NW*(K+1))        row table, padded to 16
//   v[NBUF][VB] doubles
//   acc[SB + 1] doubles          (slot SB receives the padding entries)
// DBG (measurement only): bit 0 skips the staging, bit 1 the shared-memory work.
template <int LD>
__device__ __forceinline__ uint32_t band_ld_u32(const void* p) {
    uint32_t v;
    if (LD == 0) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (LD == 1) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (LD == 2) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <int LD>
__device__ __forceinline__ double band_ld_f64(const void* p) {
    double v;
    if (LD == 0) asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (LD == 1) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (LD == 2) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// One item (all threads of the CTA call it with the same arguments). Returns
// this thread's share of the fused dot x'y (kBandRowFinal only). `smem_raw`
// is the CTA's dynamic shared memory (band_smem_bytes(T.plan) bytes, 128-byte
// aligned). The shared memory may be reused as soon as the call returns.
template <int NW, int D, int DBG = 0, int LD = 0>
__device__ __forceinline__ double band_sweep_item(const BandDev& T, const BandArgs& A, int mode,
                                                  int item, unsigned char* smem_raw,
                                                  const BandReady& ready = BandReady()) {
    const BandPlan& P = T.plan;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NBUF = P.NBUF;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + 4;
    int* s_rows = reinterpret_cast<int*>(smem_raw + 64);
    const size_t rows_bytes = (((size_t)NW * (P.K + 1) * 4) + 15) & ~(size_t)15;
    double* v_buf = reinterpret_cast<double*>(smem_raw + 64 + rows_bytes);
    double* acc_s = v_buf + (size_t)NBUF * P.VB;

    unsigned long long* trace = (DBG & 4) ? T.trace + (size_t)item * (2 * NW + 4) : nullptr;
    if ((DBG & 4) && tid == 0) trace[0] = globaltimer();
    const int sb = item / P.nparts;
    const int part = item - sb * P.nparts;
    const int seg_base = sb * P.SB;
    const int nseg = min(P.SB, P.S - seg_base);
    const int vb0 = part * P.K;
    const int nk = min(P.NVB, vb0 + P.K) - vb0;
    const int rot = sb % nk;

    // Band k of the item: first gather index and length.
    auto band_range = [&](int k, int* vbase, int* vlen) {
        int r = k + rot;
        if (r >= nk) r -= nk;
        *vbase = (vb0 + r) * P.VB;
        *vlen = min(P.VB, P.V - *vbase);
    };
    // Whole producer warp: waits until the pieces band k overlaps are ready. Every lane polls
    // its own flags (a band overlaps up to a dozen pieces; one lane polling them one after the
    // other costs an L2 round trip per piece).
    auto wait_ready = [&](int k) {
        if (ready.flags == nullptr) return;
        int vbase, vlen;
        band_range(k, &vbase, &vlen);
        const int f1 = min(ready.nflags - 1, (vbase + vlen - 1) / ready.div);
        for (int f = vbase / ready.div + lane; f <= f1; f += 32) {
            unsigned now;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];"
                             : "=r"(now)
                             : "l"(ready.flags + f)
                             : "memory");
            } while ((int)(now - ready.gen) < 0);
        }
        __syncwarp();
    };
    // Stages band k of the item into ring buffer b (producer lane only, after wait_ready).
    auto stage_band = [&](int k, int b, uint64_t keep) {
        int vbase, vlen;
        band_range(k, &vbase, &vlen);
        if (ready.flags != nullptr) {
            // what the pieces' writers stored (generic proxy) must be seen by this thread's
            // loads (no stale L1 line) and by the bulk copies
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        double* dst = v_buf + (size_t)b * P.VB;
        const unsigned bytes = (unsigned)(vlen & ~1) * 8u;
        if (vlen & 1) dst[vlen - 1] = A.v[vbase + vlen - 1];
        mbar_arrive_expect_tx(full + b, bytes);
        // at most 16 KB per copy keeps several copies in flight
        for (unsigned off = 0; off < bytes; off += 16384u) {
            const unsigned n = min(16384u, bytes - off);
            bulk_g2s(reinterpret_cast<unsigned char*>(dst) + off,
                     reinterpret_cast<const unsigned char*>(A.v + vbase) + off, n, full + b, keep);
        }
    };
    uint64_t keep = 0;
    if (warp == NW) {
        // The producer warp sets up the barriers and starts the first NBUF bands at once,
        // while the other threads clear the accumulators and fetch the row table: the first
        // band is (nearly) there when the consumers start.
        if (lane == 0) {
            for (int b = 0; b < NBUF; b++) {
                mbar_init(full + b, 1);
                mbar_init(empty + b, NW);
            }
            mbar_fence_init();
            keep = policy_evict_last();
        }
        if (!(DBG & 1))
            for (int k = 0; k < NBUF && k < nk; k++) {
                wait_ready(k);
                if (lane == 0) stage_band(k, k, keep);
            }
    } else {
        const int pcount = NW * 32;
        for (int s = tid; s <= nseg; s += pcount) acc_s[s] = 0.0;
        const int* src = T.row_ptr + (size_t)item * NW * (P.K + 1);
        for (int i = tid; i < NW * (P.K + 1); i += pcount) s_rows[i] = src[i];
    }
    __syncthreads();
    if ((DBG & 4) && tid == 0) trace[1] = globaltimer();

    if (warp == NW) {
        // ===== producer: stage the remaining bands as buffers are released =====
        if (lane > 0 && mode == kBandColScale && A.W != nullptr) {
            // pull the weights of the epilogue into L2 meanwhile
            const char* wbase = reinterpret_cast<const char*>(A.W + seg_base);
            const long long wbytes = (long long)nseg * 8;
            for (long long o = (long long)(lane - 1) * 128; o < wbytes; o += 31 * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(wbase + o));
        }
        if (!(DBG & 1) && (lane == 0 || ready.flags != nullptr)) {
            int b = 0;
            unsigned par = 0;  // parity of the previous use of buffer b
            for (int k = NBUF; k < nk; k++) {
                wait_ready(k);
                if (lane == 0) {
                    mbar_wait(empty + b, par);
                    stage_band(k, b, keep);
                }
                if (++b == NBUF) {
                    b = 0;
                    par ^= 1u;
                }
            }
        }
    } else {
        // ===== consumers =====
        // Rows [r_begin, r_end) of this warp, all tiles back to back. Two
        // register batches of D rows alternate: one is consumed while the
        // other is in flight. The stream ends with 2*D spare rows, so batch
        // loads need no bounds checks.
        const int* rp = s_rows + warp * (P.K + 1);
        const int r_begin = rp[0], r_end = rp[nk];
        const unsigned char* lane_base = T.stream + (size_t)lane * 4;
        uint32_t ka[D], kb[D];
        double aa[D], ab[D];
        auto load_batch = [&](uint32_t* kk, double* av, int r0) {
            const unsigned char* p = lane_base + (size_t)r0 * kBandRowBytes;
#pragma unroll
            for (int u = 0; u < D; u++) {
                kk[u] = band_ld_u32<LD>(p + u * kBandRowBytes);
                av[u] = band_ld_f64<LD>(p + u * kBandRowBytes + 128 + lane * 4);
            }
        };
        int k = -1, buf = -1;
        unsigned par = 1;
        int boundary = r_begin;
        const double* v_s = v_buf;
        double sum = 0.0;
        double last_v = 0.0;
        unsigned long long waited = 0;  // DBG & 4: ns spent waiting for bands
        auto advance = [&](int rr) {
            do {
                if (k >= 0) {
                    // the buffer's last gathered value must have arrived before release
                    asm volatile("" ::"d"(last_v) : "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + buf);
                }
                k++;
                if (++buf == NBUF) buf = 0;
                if (buf == 0) par ^= 1u;
                unsigned long long tw0 = 0;
                if (DBG & 4) tw0 = globaltimer();
                if (!(DBG & 1)) mbar_wait(full + buf, par);
                if (DBG & 4) waited += globaltimer() - tw0;
                v_s = v_buf + (size_t)buf * P.VB;
                boundary = rp[k + 1];
            } while (rr == boundary && k < nk - 1);
        };
        auto consume_batch = [&](const uint32_t* kk, const double* av, int r0) {
#pragma unroll
            for (int u = 0; u < D; u++) {
                const int rr = r0 + u;
                if (rr < r_end) {
                    if (rr == boundary) advance(rr);
                    const uint32_t key = kk[u];
                    const double a = av[u];
                    if (!(DBG & 2)) {
                        last_v = v_s[key & 0x7fffu];
                        sum = sum + __dmul_rn(last_v, a);
                        if (key & kBandLast) {
                            acc_s[key >> 16] += sum;
                            sum = 0.0;
                        }
                    } else {
                        sum += a + __uint_as_float(key);
                    }
                }
            }
        };
        load_batch(ka, aa, r_begin);
        for (int r = r_begin; r < r_end; r += 2 * D) {
            load_batch(kb, ab, r + D);
            consume_batch(ka, aa, r);
            load_batch(ka, aa, r + 2 * D);
            consume_batch(kb, ab, r + D);
        }
        // tiles after this warp's last row: release and (so that no bulk copy
        // is in flight when the CTA exits) wait for the remaining bands
        while (k < nk - 1) advance(-1);
        if (DBG & 2) acc_s[lane] = sum;
        if ((DBG & 4) && lane == 0) {
            trace[4 + warp] = globaltimer();
            trace[4 + NW + warp] = waited;
        }
    }
    __syncthreads();
    if ((DBG & 4) && tid == 0) trace[2] = globaltimer();

    double dot = 0.0;
    {
        // restrict: lets the (at most a handful of) iterations' loads go out together
        const double* __restrict__ Wp = A.W;
        const double* __restrict__ xp = A.x;
        const double* __restrict__ Wsp = A.Ws;
        double* __restrict__ outp = mode == kBandPartial ? T.partials + (size_t)part * P.S : A.out;
        if (mode == kBandColScale) {
#pragma unroll 4
            for (int s = tid; s < nseg; s += (NW + 1) * 32) {
                const int g = seg_base + s;
                outp[g] = Wp ? __dmul_rn(acc_s[s], Wp[g]) : acc_s[s];
            }
        } else if (mode == kBandRowFinal) {
#pragma unroll 4
            for (int s = tid; s < nseg; s += (NW + 1) * 32) {
                const int g = seg_base + s;
                const double xv = xp[g];
                const double yv = (Wsp ? __dmul_rn(xv, Wsp[g]) : 0.0) + acc_s[s];
                outp[g] = yv;
                dot += __dmul_rn(xv, yv);
            }
        } else {
            for (int s = tid; s < nseg; s += (NW + 1) * 32) outp[seg_base + s] = acc_s[s];
        }
    }
    if (tid == 0) {
        for (int b = 0; b < NBUF; b++) {
            mbar_inval(full + b);
            mbar_inval(empty + b);
        }
    }
    __syncthreads();
    if ((DBG & 4) && tid == 0) trace[3] = globaltimer();
    return dot;
}

template <int NW, int D, int DBG = 0, int LD = 0>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
band_sweep_kernel(BandDev T, BandArgs A, int mode, Reduce red, CrState* st) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_red[32];
    __shared__ int s_flag;
    if (st != nullptr && st->done) return;
    double dot = 0.0;
    for (int item = blockIdx.x; item < T.plan.nitems; item += gridDim.x)
        dot += band_sweep_item<NW, D, DBG, LD>(T, A, mode, item, smem_raw);
    if (mode == kBandRowFinal) {
        const double mine = block_sum(dot, s_red);
        double ts, ts2, tm;
        if (grid_reduce(red, mine, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) &&
            threadIdx.x == 0) {
            A.out[T.plan.S] = ts;
            // (plain, no slot): the scalar step and the time stamp belong to a later kernel
            if (st && !(A.apply_mode == kApplyPlain && A.slot == kSlotNone))
                after_apply(st, A.apply_mode, ts, A.slot);
        }
    }
}

// out = epilogue(sum over parts, in order) for sweeps that ran in partial mode.
__global__ void __launch_bounds__(kBlock)
band_combine_kernel(BandDev T, BandArgs A, int mode, Reduce red, CrState* st) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    if (st != nullptr && st->done) return;
    const int S = T.plan.S, nparts = T.plan.nparts;
    double dot = 0.0;
    for (int g = blockIdx.x * kBlock + threadIdx.x; g < S; g += gridDim.x * kBlock) {
        double acc = 0.0;
        for (int p = 0; p < nparts; p++) acc += __ldcg(T.partials + (size_t)p * S + g);
        if (mode == kBandColScale) {
            A.out[g] = A.W ? __dmul_rn(acc, A.W[g]) : acc;
        } else {
            const double xv = A.x[g];
            const double yv = (A.Ws ? __dmul_rn(xv, A.Ws[g]) : 0.0) + acc;
            A.out[g] = yv;
            dot += __dmul_rn(xv, yv);
        }
    }
    if (mode == kBandRowFinal) {
        const double mine = block_sum(dot, s_red);
        double ts, ts2, tm;
        if (grid_reduce(red, mine, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) &&
            threadIdx.x == 0) {
            A.out[S] = ts;
            if (st && !(A.apply_mode == kApplyPlain && A.slot == kSlotNone))
                after_apply(st, A.apply_mode, ts, A.slot);
        }
    }
}

// The same for many parts (>= kBandWideParts, e.g. 136 parts x 7000 rows of a transportation
// LP): a CTA takes 32 segments, its 8 warps sum every 8th part each (coalesced rows of the
// partials, several loads in flight) and warp 0 adds the 8 sums in warp order - a fixed
// order, so the result is deterministic. (band_combine_kernel walks the parts one thread per
// segment: 59 us for 7.7 MB there, latency bound on 28 CTAs.)
constexpr int kBandWideParts = 16;

__global__ void __launch_bounds__(kBlock)
band_combine_wide_kernel(BandDev T, BandArgs A, int mode, Reduce red, CrState* st) {
    static_assert(kBlock == 256, "8 warps of 32 segments");
    __shared__ double s_part[8][32];
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    if (st != nullptr && st->done) return;
    const int S = T.plan.S, nparts = T.plan.nparts;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double dot = 0.0;
    for (int g0 = blockIdx.x * 32; g0 < S; g0 += gridDim.x * 32) {
        const int g = g0 + lane;
        double acc = 0.0;
        if (g < S) {
#pragma unroll 4
            for (int p = warp; p < nparts; p += 8) acc += __ldcg(T.partials + (size_t)p * S + g);
        }
        s_part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0 && g < S) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) tot += s_part[w][lane];
            if (mode == kBandColScale) {
                A.out[g] = A.W ? __dmul_rn(tot, A.W[g]) : tot;
            } else {
                const double xv = A.x[g];
                const double yv = (A.Ws ? __dmul_rn(xv, A.Ws[g]) : 0.0) + tot;
                A.out[g] = yv;
                dot += __dmul_rn(xv, yv);
            }
        }
        __syncthreads();
    }
    if (mode == kBandRowFinal) {
        const double mine = block_sum(dot, s_red);
        double ts, ts2, tm;
        if (grid_reduce(red, mine, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) &&
            threadIdx.x == 0) {
            A.out[S] = ts;
            if (st && !(A.apply_mode == kApplyPlain && A.slot == kSlotNone))
                after_apply(st, A.apply_mode, ts, A.slot);
        }
    }
}

// ---- host side ----

inline size_t band_smem_bytes(const BandPlan& P) {
    const size_t rows_bytes = (((size_t)P.NW * (P.K + 1) * 4) + 15) & ~(size_t)15;
    return 64 + rows_bytes + (size_t)P.NBUF * P.VB * 8 + ((size_t)P.SB + 1) * 8;
}

// Fills the derived fields of a plan from (V, S, VB, SB, K, NW, NBUF).
inline bool band_finish_plan(BandPlan* P) {
    if (P->V <= 0 || P->S <= 0 || P->VB <= 0 || P->SB <= 0) return false;
    if (P->VB > kBandMaxVB || (P->VB & 1) || P->SB > 65535) return false;
    P->NVB = (P->V + P->VB - 1) / P->VB;
    P->NSB = (P->S + P->SB - 1) / P->SB;
    if (P->K <= 0 || P->K > P->NVB) P->K = P->NVB;
    if (P->K > kBandMaxSteps) return false;
    P->nparts = (P->NVB + P->K - 1) / P->K;
    P->nitems = P->NSB * P->nparts;
    P->smem = band_smem_bytes(*P);
    return P->smem <= 227 * 1024;
}

// Runs fn(begin, end) over [0, n) on up to `threads` host threads (contiguous ranges).
template <class F>
inline void band_parallel_for(size_t n, int threads, F fn) {
    threads = (int)std::min<size_t>((size_t)std::max(1, threads), std::max<size_t>(1, n));
    if (threads <= 1) {
        fn((size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++) {
        const size_t b = n * k / threads, e = n * (k + 1) / threads;
        pool.emplace_back([=] { fn(b, e); });
    }
    for (std::thread& t : pool) t.join();
}

inline int band_build_threads() {
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(8u, std::max(1u, hw / 2));  // two sweeps are built side by side
}

// Re-tiles a compressed structure (segments ptr[0..S], gather indices idx in
// [0,V), ascending per segment) into the banded row streams. Returns false when
// a run is longer than max_run (such structures suit the generic sweep).
//
// `spilled` (optional): instead of refusing, segments with such a run are left out of the
// streams altogether (their accumulators stay 0) and listed, ascending; the caller sweeps them
// with the generic kernel (dense rows / columns of an otherwise well-spread matrix). Refuses
// when the spilled segments hold more than 3/4 of the entries.
inline bool band_build(const BandPlan& P, const int* ptr, const int* idx, const double* val,
                       BandHost* H, int max_run = kBandMaxRun, std::vector<int>* spilled = nullptr) {
    const int S = P.S, VB = P.VB, SB = P.SB, NVB = P.NVB, NSB = P.NSB, NW = P.NW;
    const size_t ntiles = (size_t)NSB * NVB;
    std::vector<char> skip;
    if (spilled) {
        spilled->clear();
        long long spilled_entries = 0;
        for (int s = 0; s < S; s++) {
            int p = ptr[s];
            const int pe = ptr[s + 1];
            bool is_long = false;
            while (p < pe && !is_long) {
                const int vb = idx[p] / VB;
                int q = p + 1;
                while (q < pe && idx[q] / VB == vb) q++;
                is_long = q - p > max_run;
                p = q;
            }
            if (is_long) {
                if (skip.empty()) skip.assign((size_t)S, 0);
                skip[s] = 1;
                spilled->push_back(s);
                spilled_entries += pe - ptr[s];
            }
        }
        if (4 * spilled_entries > 3 * (long long)P.nnz) return false;
    }
    const bool any_skip = !skip.empty();
    struct Run {
        int first;            // first entry in idx/val
        unsigned short seg;   // local segment
        unsigned short len;
    };
    // pass 1: runs per (tile, warp). A segment belongs to the same warp in
    // every tile of its block (local segment % NW), so all additions to one
    // accumulator are issued by one warp, in program order; warps never have
    // to synchronise with each other between tiles.
    const size_t ntw = ntiles * (size_t)NW;
    std::vector<long long> tcount(ntw + 1, 0);
    const int nthreads = band_build_threads();
    // The tiles of different segment blocks are disjoint: blocks are counted (and below
    // scattered) side by side; inside a tile the runs keep their segment order.
    std::vector<char> too_long((size_t)NSB, 0);
    band_parallel_for((size_t)NSB, nthreads, [&](size_t sb_begin, size_t sb_end) {
        for (size_t sb = sb_begin; sb < sb_end; sb++) {
            const int s_end = (int)std::min<long long>(S, (long long)(sb + 1) * SB);
            for (int s = (int)sb * SB; s < s_end; s++) {
                if (any_skip && skip[s]) continue;
                const size_t trow = sb * NVB;
                const int w = (s % SB) % NW;
                int p = ptr[s];
                const int pe = ptr[s + 1];
                while (p < pe) {
                    const int vb = idx[p] / VB;
                    int q = p + 1;
                    while (q < pe && idx[q] / VB == vb) q++;
                    if (q - p > max_run) too_long[sb] = 1;
                    tcount[(trow + vb) * NW + w + 1]++;
                    p = q;
                }
            }
        }
    });
    for (char f : too_long)
        if (f) return false;
    for (size_t t = 0; t < ntw; t++) tcount[t + 1] += tcount[t];
    std::vector<Run> runs((size_t)tcount[ntw]);
    {
        std::vector<long long> next(tcount.begin(), tcount.end() - 1);
        band_parallel_for((size_t)NSB, nthreads, [&](size_t sb_begin, size_t sb_end) {
            for (size_t sb = sb_begin; sb < sb_end; sb++) {
                const int s_end = (int)std::min<long long>(S, (long long)(sb + 1) * SB);
                for (int s = (int)sb * SB; s < s_end; s++) {
                    if (any_skip && skip[s]) continue;
                    const size_t trow = sb * NVB;
                    const int w = (s % SB) % NW;
                    int p = ptr[s];
                    const int pe = ptr[s + 1];
                    while (p < pe) {
                        const int vb = idx[p] / VB;
                        int q = p + 1;
                        while (q < pe && idx[q] / VB == vb) q++;
                        runs[(size_t)next[(trow + vb) * NW + w]++] =
                            Run{p, (unsigned short)(s % SB), (unsigned short)(q - p)};
                        p = q;
                    }
                }
            }
        });
    }
    // pass 2: deal the runs of every (tile, warp) to the 32 lanes. place[r] = lane << 16 | offset.
    //
    // Runs longer than one entry go first, longest first, each to the least loaded lane (ties:
    // lowest lane), so loads differ by <= 1 whenever there are enough short runs. The single
    // entries - nine out of ten on a matrix without structure - then fill the rows one by one,
    // and which entry goes to which lane is chosen by shared-memory bank: the kernel gathers
    // v[index] (LDS.64) and adds into acc[segment] (LDS.64 + STS.64) once per entry, a 64-bit
    // access of a warp is served half-warp by half-warp, and a half-warp takes as many wavefronts
    // as its most loaded bank pair (index mod 16, segment mod 16). Dealt at random that is ~2.9
    // and ~1.8 per access (13 wavefronts per row, measured: tools/band_conflicts.cu, ncu); here
    // every half-row gets 16 entries with distinct gather banks and, where the pool allows,
    // distinct accumulator banks. (bank_aware = false keeps the plain dealing for comparison.)
    std::vector<uint32_t> place(runs.size());
    std::vector<int> trows(ntw, 0);  // rows of (tile, warp)
    const bool bank_aware = [] {
        const char* env = std::getenv("IPXGPU_BAND_DEAL");
        return !(env && std::string(env) == "plain");
    }();
    band_parallel_for(ntw, nthreads, [&](size_t t_begin, size_t t_end) {
        std::vector<long long> order;
        std::vector<int> bucket_start;
        std::vector<long long> pool[16];       // single entries by gather bank pair
        std::vector<uint16_t> used_g, used_a;  // per (row, half): bank pairs taken
        for (size_t t = t_begin; t < t_end; t++) {
            const long long r0 = tcount[t], r1 = tcount[t + 1];
            const int nr = (int)(r1 - r0);
            if (nr == 0) continue;
            const int vb = (int)((t / NW) % NVB);
            int maxlen = 0;
            long long total = 0;
            for (long long r = r0; r < r1; r++) {
                maxlen = std::max(maxlen, (int)runs[r].len);
                total += runs[r].len;
            }
            int load[32];
            for (int l = 0; l < 32; l++) load[l] = 0;
            {
                bucket_start.assign(maxlen + 2, 0);
                for (long long r = r0; r < r1; r++) bucket_start[maxlen - runs[r].len + 1]++;
                for (int b = 0; b <= maxlen; b++) bucket_start[b + 1] += bucket_start[b];
                order.resize(nr);
                for (long long r = r0; r < r1; r++)
                    order[bucket_start[maxlen - runs[r].len]++] = r;
            }
            if (!bank_aware) {
                for (int k = 0; k < nr; k++) {
                    const long long r = order[k];
                    int best = 0;
                    for (int l = 1; l < 32; l++)
                        if (load[l] < load[best]) best = l;
                    place[r] = ((uint32_t)best << 16) | (uint32_t)load[best];
                    load[best] += runs[r].len;
                }
            } else {
                // Row by row, every lane that is free takes a new run: the longest one in the
                // pool (buckets by the gather bank of the run's first entry, longest at the
                // back) whose entries find their gather banks free in the rows it will occupy
                // and whose accumulator bank is free in the row where it ends.
                const int cap = (int)((total + 31) / 32) + 2 * maxlen + 2;
                used_g.assign((size_t)cap * 2, 0);
                used_a.assign((size_t)cap * 2, 0);
                for (int b = 0; b < 16; b++) pool[b].clear();
                for (int k = nr - 1; k >= 0; k--) {  // order[] is longest first
                    const long long r = order[k];
                    pool[(idx[runs[r].first] - vb * VB) & 15].push_back(r);
                }
                auto fits = [&](long long r, int half, int row) {
                    const Run& R = runs[r];
                    if (row + R.len > cap) return false;
                    for (int j = 0; j < R.len; j++)
                        if (used_g[(size_t)(row + j) * 2 + half] &
                            (1u << ((idx[R.first + j] - vb * VB) & 15)))
                            return false;
                    return !(used_a[(size_t)(row + R.len - 1) * 2 + half] & (1u << (R.seg & 15)));
                };
                long long left = nr;
                for (int row = 0; left > 0; row++) {
                    for (int l = 0; l < 32 && left > 0; l++) {
                        if (load[l] != row) continue;
                        const int half = l >> 4;
                        int bb = -1, bpos = -1, blen = 0;
                        size_t bsize = 0;
                        int fb = -1, flen = 0;  // fallback: longest run at the back of a bucket
                        for (int b = 0; b < 16; b++) {
                            const size_t sz = pool[b].size();
                            if (sz == 0) continue;
                            const int backlen = runs[pool[b][sz - 1]].len;
                            if (backlen > flen) {
                                flen = backlen;
                                fb = b;
                            }
                            const int look = (int)std::min<size_t>(sz, 10);
                            for (int q = 0; q < look; q++) {
                                const long long r = pool[b][sz - 1 - q];
                                const int len = runs[r].len;
                                if (len < blen || (len == blen && sz <= bsize)) break;
                                if (fits(r, half, row)) {
                                    bb = b;
                                    bpos = (int)(sz - 1 - q);
                                    blen = len;
                                    bsize = sz;
                                    break;
                                }
                            }
                        }
                        int b = bb, pos = bpos;
                        if (b < 0) {
                            b = fb;
                            pos = (int)pool[b].size() - 1;
                        }
                        const long long r = pool[b][pos];
                        pool[b].erase(pool[b].begin() + pos);  // keeps the bucket sorted by length
                        left--;
                        const Run& R = runs[r];
                        place[r] = ((uint32_t)l << 16) | (uint32_t)row;
                        for (int j = 0; j < R.len && row + j < cap; j++)
                            used_g[(size_t)(row + j) * 2 + half] |=
                                (uint16_t)(1u << ((idx[R.first + j] - vb * VB) & 15));
                        if (row + R.len - 1 < cap)
                            used_a[(size_t)(row + R.len - 1) * 2 + half] |=
                                (uint16_t)(1u << (R.seg & 15));
                        load[l] = row + R.len;
                    }
                }
            }
            int mx = 0;
            for (int l = 0; l < 32; l++) mx = std::max(mx, load[l]);
            trows[t] = mx;
        }
    });
    // pass 3: row offsets in stream order (item, warp, step)
    const int K = P.K, nparts = P.nparts;
    H->row_ptr.assign((size_t)P.nitems * NW * (K + 1), 0);
    std::vector<long long> tile_warp_row(ntiles * (size_t)NW, 0);
    long long rows = 0;
    for (int sb = 0; sb < NSB; sb++) {
        for (int part = 0; part < nparts; part++) {
            const int item = sb * nparts + part;
            const int vb0 = part * K;
            const int nk = std::min(NVB, vb0 + K) - vb0;
            const int rot = sb % nk;
            for (int w = 0; w < NW; w++) {
                int* rp = H->row_ptr.data() + ((size_t)item * NW + w) * (K + 1);
                for (int k = 0; k < nk; k++) {
                    int r = k + rot;
                    if (r >= nk) r -= nk;
                    const size_t t = (size_t)sb * NVB + vb0 + r;
                    rp[k] = (int)rows;
                    tile_warp_row[t * NW + w] = rows;
                    rows += trows[t * NW + w];
                }
                for (int k = nk; k <= K; k++) rp[k] = (int)rows;
                if (rows >= (long long)INT32_MAX / 2) return false;
            }
        }
    }
    H->rows = rows;
    const long long rows_alloc = rows + kBandTailRows;  // spare rows: unchecked batch loads
    // pass 4: fill; padding = (segment SB, last, index 0, value 0)
    H->stream.assign((size_t)rows_alloc * 96, 0u);
    {
        const uint32_t padkey = ((uint32_t)SB << 16) | kBandLast;
        uint32_t* base = H->stream.data();
        band_parallel_for((size_t)rows, nthreads, [=](size_t r_begin, size_t r_end) {
            for (size_t r = r_begin; r < r_end; r++) {
                uint32_t* row = base + r * 96;
                for (int l = 0; l < 32; l++) row[l] = padkey;
            }
        });
    }
    // every (tile, warp) owns its rows: the tiles are filled independently
    band_parallel_for(ntw, nthreads, [&](size_t t_begin, size_t t_end) {
        for (size_t t = t_begin; t < t_end; t++) {
            const int vb = (int)((t / NW) % NVB);
            for (long long r = tcount[t]; r < tcount[t + 1]; r++) {
                const Run& R = runs[r];
                const int l = (int)(place[r] >> 16), off = (int)(place[r] & 0xffffu);
                const long long row0 = tile_warp_row[t] + off;
                for (int j = 0; j < R.len; j++) {
                    uint32_t* row = H->stream.data() + (size_t)(row0 + j) * 96;
                    uint32_t key = ((uint32_t)R.seg << 16) | (uint32_t)(idx[R.first + j] - vb * VB);
                    if (j == R.len - 1) key |= kBandLast;
                    row[l] = key;
                    reinterpret_cast<double*>(row + 32)[l] = val[R.first + j];
                }
            }
        }
    });
    long long filled = 0;
    for (const Run& R : runs) filled += R.len;
    H->pad_entries = rows * 32 - filled;
    return true;
}

}  // namespace ipxgpu
