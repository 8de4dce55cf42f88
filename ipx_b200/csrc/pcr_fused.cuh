// The whole (preconditioned) Conjugate Residuals solve with C = A*W*A' as ONE
// persistent cooperative kernel (reference src/conjugate_residuals.cc:14-88,
// :90-213 with C = NormalMatrix, P = DiagonalPrecond).
//
// The launch-per-stage loop of cr_kernels.cuh pays ~7-10 us of launch, fill
// and drain per kernel, 5 kernels per CR iteration, against ~40 us of memory
// traffic. Here one CTA per SM stays resident for the whole solve and the
// stages of an iteration are separated by grid barriers:
//
//   update     y += a p; r -= a Cp; [s -= a q]      own slice of the m-vectors
//   -- barrier (s complete: sweep 1 gathers it)
//   sweep 1    t = W .* (A' s)                      banded items over columns
//   -- barrier
//   sweep 2    partial[part] = A[:, part] t         banded items over rows
//   -- barrier
//   combine    Cs = Ws .* s + sum(partials); dot    own slice
//   -- barrier (dot -> beta)
//   direction  p = s + b p; Cp = Cs + b Cp; [q = Cp./diag]; pdot; [s = r./diag]
//   -- barrier (pdot, resnorm -> tests, alpha)
//
// Every CTA owns a fixed slice of the m-vectors and reduces the per-CTA
// partials of a stage in CTA order while it waits at the barrier, so all CTAs hold
// identical scalars and take identical decisions; results are run-to-run
// deterministic. The scalar logic is the reference's, in its order.
#pragma once

#include "band_sweep.cuh"
#include "cr_kernels.cuh"

namespace ipxgpu {

// Grid-wide synchronisation of the cooperative grid: one monotonically
// increasing arrival counter (never reset: the k-th barrier completes when it
// reaches k * gridDim.x) and, per stage, an array of per-CTA values that every
// CTA reduces in CTA order after the barrier, so all CTAs obtain identical
// results.
struct GridSync {
    unsigned* count;  // arrival counter, zero before the launch
    double* vals;     // [kFusedStages][3][gridDim.x]
};

struct FusedArgs {
    BandDev T1, T2;
    CrVectors v;
    const double* rhs;
    const double* Wc;   // structural weights or nullptr (1)
    const double* Ws;   // slack weights or nullptr (0)
    double* t;          // intermediate n-vector
    int zero_start;
    GridSync sync;
    double* abort_word; // set nonzero by CTA 0 when the host asked to stop
    unsigned* block_tickets;  // [T2.plan.NSB], zero before the launch: items done per row block
    // Readiness flags, zero before the launch (nullptr: grid barriers instead). t_ready[i]: the
    // number of applies whose sweep-1 item i has written its block of t; x_ready[c]: the number
    // of updates after which CTA c's slice of the vector the next apply gathers is final.
    unsigned* t_ready;
    unsigned* x_ready;
    CrState* st;
    const volatile int* abort_flag;  // host-mapped; nonzero asks the solve to stop
    unsigned long long* trace;       // tuning only: globaltimer of CTA 0 at every stage end
    int trace_cap;
    // Sharded contexts: every rank's exchange buffer (as mapped in this process),
    // y[2][xmpad] doubles followed by flags[nranks][gridDim.x]. nranks == 1: unused.
    int nranks, rank;
    double* const* peers;
    size_t xmpad;
    unsigned xgen_base;              // cross-GPU synchronisations before this solve
    // Push exchange (default): records[2][nranks][xmpad] of 16 bytes at byte offset xll_off of
    // every rank's buffer, followed by final[2][xmpad]; 0 = pull exchange (flags + P2P loads).
    size_t xll_off;
    int xtwo_phase;  // reduce-scatter + all-gather of records instead of all-to-all
    int xfold;       // sharded: exchange a row block's rows right after its sweep-2 items (no grid barrier)
};

// 16-byte record of two self-validating 64-bit words {generation | half of the value}: 64-bit
// words arrive atomically over NVLink (the basis of NCCL's LL protocol), so a reader that
// finds the generation in both words holds the value - no flag, no fence.sys, no round trip.
__device__ __forceinline__ void xll_store(ulonglong2* p, unsigned gen, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long w0 = ((unsigned long long)gen << 32) | (bits & 0xffffffffull);
    const unsigned long long w1 = ((unsigned long long)gen << 32) | (bits >> 32);
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1)
                 : "memory");
}
__device__ __forceinline__ bool xll_load(const ulonglong2* p, unsigned gen, double* val) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];"
                 : "=l"(w0), "=l"(w1)
                 : "l"(p)
                 : "memory");
    *val = __longlong_as_double((long long)(((w1 & 0xffffffffull) << 32) | (w0 & 0xffffffffull)));
    return (unsigned)(w0 >> 32) == gen && (unsigned)(w1 >> 32) == gen;
}

// Waits for a record of generation gen; a peer that never arrives must not hang this GPU: after
// 2^24 polls (or when another thread gave up) the abort word is raised and the wait ends.
__device__ __forceinline__ void xchg_wait(const ulonglong2* p, unsigned gen, double* val,
                                       double* abort_word) {
    long long spins = 0;
    while (!xll_load(p, gen, val)) {
        if ((++spins & 0xffff) == 0) {
            if (spins > (1ll << 24) || __ldcg(abort_word) != 0.0) {
                *abort_word = 2.0;
                break;
            }
        }
    }
}

constexpr int kFusedStages = 6;
constexpr int kMaxGridSync = 256;  // CTAs of the persistent kernel (one per SM)
enum FusedStage : int { kStSweep1 = 0, kStSweep2, kStCombine, kStInit, kStDirection, kStUpdate };

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async;" ::: "memory");
}

// Arrives at barrier number `gen` (1, 2, ...) and waits for it. Generic-proxy
// writes of any CTA before its arrival are visible to generic and bulk-copy
// reads of every CTA after the wait. With stage >= 0, thread 0's (v0, v1, v2)
// are published and the sums / sum / max over all CTAs are returned to all
// threads.
__device__ __forceinline__ void grid_sync(const GridSync& g, unsigned gen, int stage, double v0,
                                          double v1, double v2, double* s_bcast, double* out0,
                                          double* out1, double* out2) {
    const int nblk = gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (stage >= 0) {
            double* base = g.vals + (size_t)stage * 3 * nblk;
            base[blockIdx.x] = v0;
            base[nblk + blockIdx.x] = v1;
            base[2 * nblk + blockIdx.x] = v2;
        }
        // Global memory is coherent in L2 for both proxies; the proxy fences of
        // thread 0 order its release / acquire against the bulk copies that
        // read what other CTAs wrote before the barrier.
        fence_proxy_async();
        __threadfence();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(g.count) : "memory");
        const unsigned target = gen * (unsigned)nblk;
        unsigned now;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(g.count) : "memory");
        } while ((int)(now - target) < 0);
        __threadfence();  // acquire; also drops this SM's stale L1 lines
        fence_proxy_async();
    }
    __syncthreads();
    if (stage >= 0) {
        // Every CTA reduces the published values in CTA order (identical results everywhere).
        // The loads go out together (one thread per value), the sums run in a fixed order.
        const double* base = g.vals + (size_t)stage * 3 * nblk;
        double* s_vals = s_bcast + 4;  // [3 * kMaxGridSync]
        for (int q = threadIdx.x; q < 3 * nblk; q += blockDim.x) s_vals[q] = __ldcg(base + q);
        __syncthreads();
        if (threadIdx.x < 32) {
            double s = 0.0, s2 = 0.0, mx = 0.0;
            for (int b = threadIdx.x; b < nblk; b += 32) {
                s += s_vals[b];
                s2 += s_vals[nblk + b];
                mx = fmax(mx, s_vals[2 * nblk + b]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_down_sync(0xffffffffu, s, o);
                s2 += __shfl_down_sync(0xffffffffu, s2, o);
                mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
            }
            if (threadIdx.x == 0) {
                s_bcast[0] = s;
                s_bcast[1] = s2;
                s_bcast[2] = mx;
            }
        }
        __syncthreads();
        *out0 = s_bcast[0];
        *out1 = s_bcast[1];
        *out2 = s_bcast[2];
    }
}

// CTA-wide sum, sum, max in one pass; valid in thread 0. s_red: 96 doubles.
__device__ __forceinline__ void fused_block_reduce(double* s_red, double& a, double& b, double& c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
        c = fmax(c, __shfl_down_sync(0xffffffffu, c, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) {
        s_red[warp] = a;
        s_red[32 + warp] = b;
        s_red[64 + warp] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = s_red[0];
        b = s_red[32];
        c = s_red[64];
        for (int w = 1; w < nw; w++) {
            a += s_red[w];
            b += s_red[32 + w];
            c = fmax(c, s_red[64 + w]);
        }
    }
}

// SH = 0: the single-shard kernel (no exchange code compiled in: the 64-register budget of a
// 1024-thread CTA is tight); SH = 1: column shards.
template <int NW, int D, int DBG = 0, int SH = 0>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
pcr_fused_kernel(FusedArgs F) {
    const int nranks = SH ? F.nranks : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_red[96];
    __shared__ double s_bcast[4 + 3 * kMaxGridSync];
    __shared__ CrState s_st;   // this CTA's replica of the solve state

    const int tid = threadIdx.x;
    const int nthr = (NW + 1) * 32;
    const CrVectors& v = F.v;
    const int m = v.m;
    const int chunk = (m + gridDim.x - 1) / gridDim.x;
    const int i0 = min(m, (int)blockIdx.x * chunk), i1 = min(m, i0 + chunk);
    const bool lead = blockIdx.x == 0 && tid == 0;

    if (tid == 0) {
        s_st = *F.st;
        s_st.t_last = globaltimer();
    }
    __syncthreads();
    const bool precond = s_st.precond != 0;

    auto stamp_lead = [&](int slot) {
        if (lead) stamp(&s_st, slot);
    };
    int ntrace = 0;
    auto trace = [&]() {
        if (lead && F.trace && ntrace < F.trace_cap) F.trace[ntrace++] = globaltimer();
    };
    unsigned gen = 0;  // barriers passed so far (identical in all CTAs)
    int s_applies = 0; // C.Apply calls so far (identical in all CTAs and ranks)
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
    // Barrier with reduction; (a, b, c) are this thread's shares of sum, sum,
    // max. Totals land in r0, r1, r2 of every thread.
    auto sync_stage = [&](int stage, double a, double b, double c2) {
        fused_block_reduce(s_red, a, b, c2);
        trace();
        grid_sync(F.sync, ++gen, stage, a, b, c2, s_bcast, &r0, &r1, &r2);
        trace();
    };
    auto sync_plain = [&]() {
        trace();
        grid_sync(F.sync, ++gen, -1, 0.0, 0.0, 0.0, s_bcast, &r0, &r1, &r2);
        trace();
    };

    // With readiness flags (single shard, sweep 2 folded) a CTA goes from one stage to the next
    // as soon as the pieces IT needs are there: sweep 1 waits per band of x for the slices of
    // the CTAs that updated it, sweep 2 per band of t for the sweep-1 items that wrote it. A CTA
    // that is ahead no longer waits for the slowest one at two of the four grid barriers of an
    // iteration (the SMs' streaming rates differ by 10-15 %), it starts its next item instead.
    const bool flags_on = F.t_ready != nullptr && nranks == 1 &&
                          F.T2.plan.nitems <= (int)gridDim.x;
    unsigned x_gen = 0;  // updates so far (identical in all CTAs)

    // lhs (own m-slice) = A W A' x; returns the total x'lhs to all threads. x_flagged: x was
    // written by the update stage and is guarded by x_ready (no grid barrier since);
    // extra_max: rides the reduction that ends the apply (the update's residual norm).
    auto apply = [&](const double* x, double* lhs, bool x_flagged, double extra_max) __attribute__((always_inline)) -> double {
        BandArgs a1{x, F.Wc, nullptr, nullptr, F.t, kApplyPlain, kSlotNone};
        BandReady rx;
        if (flags_on && x_flagged) {
            rx.flags = F.x_ready;
            rx.gen = x_gen;
            rx.div = chunk;
            rx.nflags = (int)gridDim.x;
        }
        for (int item = blockIdx.x; item < F.T1.plan.nitems; item += gridDim.x) {
            band_sweep_item<NW, D, DBG>(F.T1, a1, kBandColScale, item, smem_raw, rx);
            if (flags_on && tid == 0) {
                // this item's block of t is written (the item ends with a CTA barrier)
                fence_proxy_async();
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(F.t_ready + item),
                             "r"((unsigned)(s_applies + 1))
                             : "memory");
            }
        }
        if (!flags_on) sync_plain();
        BandArgs a2{F.t, nullptr, nullptr, nullptr, nullptr, kApplyPlain, kSlotNone};
        BandReady rt;
        if (flags_on) {
            rt.flags = F.t_ready;
            rt.gen = (unsigned)(s_applies + 1);
            rt.div = F.T1.plan.SB;
            rt.nflags = F.T1.plan.nitems;
        }
        const int nparts = F.T2.plan.nparts;
        double dot = 0.0;
        const bool fold_sharded = nranks > 1 && F.xfold && F.xll_off != 0 &&
                                  F.T2.plan.nitems <= (int)gridDim.x;
        if ((nranks == 1 || fold_sharded) && F.T2.plan.nitems <= (int)gridDim.x) {
            // The nparts items of a row block (all resident: one item per CTA)
            // meet at the block's counter when their partials are written; each
            // then combines its share of the block's rows, in part order. No
            // separate combine stage, one grid barrier less.
            //
            // Column shards (folded exchange): the share's rows are exchanged right there,
            // while the items of other row blocks are still sweeping. Records are addressed by
            // row, so the ranks need not cut their row blocks alike. One hop (2-3 ranks):
            // every rank gets every rank's partial and sums them in rank order. Two hops (4 and
            // more): row i belongs to rank i / ceil(m / nranks), which sums the ranks' partials
            // in rank order and pushes the final value to everybody.
            ++s_applies;
            if ((int)blockIdx.x < F.T2.plan.nitems) {
                const int item = blockIdx.x;
                band_sweep_item<NW, D, DBG>(F.T2, a2, kBandPartial, item, smem_raw, rt);
                const int sb = item / nparts, part = item - sb * nparts;
                if (tid == 0) {
                    __threadfence();
                    unsigned* cnt = F.block_tickets + sb;
                    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
                    const unsigned target = (unsigned)s_applies * (unsigned)nparts;
                    unsigned now;
                    do {
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(cnt) : "memory");
                    } while ((int)(now - target) < 0);
                    __threadfence();
                }
                __syncthreads();
                const int rb = sb * F.T2.plan.SB, re = min(m, rb + F.T2.plan.SB);
                const int share = (re - rb + nparts - 1) / nparts;
                const int q0 = rb + part * share, q1 = min(re, q0 + share);
                if (!fold_sharded) {
                    for (int i = q0 + tid; i < q1; i += nthr) {
                        double acc = 0.0;
#pragma unroll 8
                        for (int p = 0; p < nparts; p++) acc += __ldcg(F.T2.partials + (size_t)p * m + i);
                        const double xv = __ldcg(x + i);
                        const double yv = (F.Ws ? __dmul_rn(xv, F.Ws[i]) : 0.0) + acc;
                        lhs[i] = yv;
                        dot += __dmul_rn(xv, yv);
                    }
                } else {
                    const unsigned gen = F.xgen_base + (unsigned)s_applies;
                    const size_t par = (size_t)(gen & 1u) * (size_t)nranks * F.xmpad;
                    const size_t fin_off = F.xll_off + 2 * (size_t)nranks * F.xmpad * 16 +
                                           (size_t)(gen & 1u) * F.xmpad * 16;
                    const int rows_per_rank = (m + nranks - 1) / nranks;
                    char* const own = reinterpret_cast<char*>(F.peers[F.rank]);
                    const ulonglong2* own_rec = reinterpret_cast<const ulonglong2*>(own + F.xll_off) + par;
                    const ulonglong2* fin = reinterpret_cast<const ulonglong2*>(own + fin_off);
                    // pass 1: this rank's partial of every row, pushed
                    for (int i = q0 + tid; i < q1; i += nthr) {
                        double acc = 0.0;
#pragma unroll 8
                        for (int p = 0; p < nparts; p++) acc += __ldcg(F.T2.partials + (size_t)p * m + i);
                        const double mine = (F.Ws ? __dmul_rn(__ldcg(x + i), F.Ws[i]) : 0.0) + acc;
                        const int r0x = F.xtwo_phase ? i / rows_per_rank : 0;
                        const int r1x = F.xtwo_phase ? r0x + 1 : nranks;
                        for (int r = r0x; r < r1x; r++)
                            xll_store(reinterpret_cast<ulonglong2*>(
                                          reinterpret_cast<char*>(F.peers[r]) + F.xll_off) +
                                          par + (size_t)F.rank * F.xmpad + i, gen, mine);
                    }
                    // pass 2: sum in rank order - every row (one hop) or the rows this rank owns,
                    // whose finals go to everybody (two hops)
                    for (int i = q0 + tid; i < q1; i += nthr) {
                        if (F.xtwo_phase && i / rows_per_rank != F.rank) continue;
                        double tot = 0.0;
                        for (int r = 0; r < nranks; r++) {
                            double part_r = 0.0;
                            xchg_wait(own_rec + (size_t)r * F.xmpad + i, gen, &part_r, F.abort_word);
                            tot += part_r;
                        }
                        if (F.xtwo_phase) {
                            for (int r = 0; r < nranks; r++)
                                xll_store(reinterpret_cast<ulonglong2*>(
                                              reinterpret_cast<char*>(F.peers[r]) + fin_off) + i, gen, tot);
                        } else {
                            lhs[i] = tot;
                            dot += __dmul_rn(__ldcg(x + i), tot);
                        }
                    }
                    // pass 3 (two hops): the finals
                    if (F.xtwo_phase)
                        for (int i = q0 + tid; i < q1; i += nthr) {
                            double yv = 0.0;
                            xchg_wait(fin + i, gen, &yv, F.abort_word);
                            lhs[i] = yv;
                            dot += __dmul_rn(__ldcg(x + i), yv);
                        }
                }
            }
        } else {
            ++s_applies;
            for (int item = blockIdx.x; item < F.T2.plan.nitems; item += gridDim.x)
                band_sweep_item<NW, D, DBG>(F.T2, a2, kBandPartial, item, smem_raw);
            sync_plain();
        }
        if (fold_sharded) {
            // (rows exchanged above)
        } else if (nranks == 1 && F.T2.plan.nitems > (int)gridDim.x) {
            for (int i = i0 + tid; i < i1; i += nthr) {
                double acc = 0.0;
#pragma unroll 8
                for (int p = 0; p < nparts; p++) acc += __ldcg(F.T2.partials + (size_t)p * m + i);
                const double xv = x[i];
                const double yv = (F.Ws ? __dmul_rn(xv, F.Ws[i]) : 0.0) + acc;
                lhs[i] = yv;
                dot += __dmul_rn(xv, yv);
            }
        } else if (nranks > 1 && F.xll_off != 0 && F.xtwo_phase) {
            // Column shards, two-phase push exchange (4 and more ranks): the slice of CTA c is
            // cut into one piece per rank. Every rank pushes its partial products of piece r
            // to rank r only; rank r sums the ranks' records of its piece in rank order and
            // pushes the FINAL values to every rank. Per rank and exchange 2(N-1)/N*m records
            // cross NVLink instead of (N-1)*m (5.6 MB -> 1.4 MB of payload at 8 ranks), at the
            // price of a second one-way hop; all ranks read the same final values, so they
            // stay bit-identical. Records validate themselves: no flags, fences or barriers;
            // the three steps are chained per thread.
            const unsigned gen = F.xgen_base + (unsigned)s_applies;
            const size_t par = (size_t)(gen & 1u) * (size_t)nranks * F.xmpad;
            const size_t fin_off = F.xll_off + 2 * (size_t)nranks * F.xmpad * 16 +
                                   (size_t)(gen & 1u) * F.xmpad * 16;
            const int per = (i1 - i0 + nranks - 1) / nranks;
            auto wait_rec = [&](const ulonglong2* p, double* val) {
                long long spins = 0;
                while (!xll_load(p, gen, val)) {
                    if ((++spins & 0xffff) == 0) {
                        // a peer that never arrives must not hang this GPU
                        if (spins > (1ll << 24) || __ldcg(F.abort_word) != 0.0) {
                            *F.abort_word = 2.0;
                            break;
                        }
                    }
                }
            };
            for (int i = i0 + tid; i < i1; i += nthr) {
                double acc = 0.0;
#pragma unroll 8
                for (int p = 0; p < nparts; p++) acc += __ldcg(F.T2.partials + (size_t)p * m + i);
                const double mine = (F.Ws ? __dmul_rn(x[i], F.Ws[i]) : 0.0) + acc;
                const int owner = (i - i0) / per;
                ulonglong2* dst = reinterpret_cast<ulonglong2*>(
                                      reinterpret_cast<char*>(F.peers[owner]) + F.xll_off) +
                                  par + (size_t)F.rank * F.xmpad + i;
                xll_store(dst, gen, mine);
                if (owner == F.rank) {
                    const ulonglong2* rec = reinterpret_cast<const ulonglong2*>(
                                                reinterpret_cast<const char*>(F.peers[F.rank]) +
                                                F.xll_off) + par + i;
                    double tot = 0.0;
                    for (int r = 0; r < nranks; r++) {
                        double part;
                        wait_rec(rec + (size_t)r * F.xmpad, &part);
                        tot += part;
                    }
                    for (int r = 0; r < nranks; r++)
                        xll_store(reinterpret_cast<ulonglong2*>(
                                      reinterpret_cast<char*>(F.peers[r]) + fin_off) + i, gen, tot);
                }
            }
            const ulonglong2* fin = reinterpret_cast<const ulonglong2*>(
                reinterpret_cast<const char*>(F.peers[F.rank]) + fin_off);
            for (int i = i0 + tid; i < i1; i += nthr) {
                double yv;
                wait_rec(fin + i, &yv);
                lhs[i] = yv;
                dot += __dmul_rn(x[i], yv);
            }
        } else if (nranks > 1 && F.xll_off != 0) {
            // Column shards, push exchange: every rank writes its partial product of the slice
            // straight into every rank's record buffer (posted NVLink stores), then sums the
            // ranks' records of the slice from its OWN memory in rank order - bit-identical on
            // all ranks. The records validate themselves (xll_store), so there is no flag, no
            // fence.sys and no load round trip over NVLink; buffers alternate by generation
            // parity (a rank can be at most one exchange ahead of a peer).
            const unsigned gen = F.xgen_base + (unsigned)s_applies;
            const size_t par = (size_t)(gen & 1u) * (size_t)nranks * F.xmpad;
            for (int i = i0 + tid; i < i1; i += nthr) {
                double acc = 0.0;
#pragma unroll 8
                for (int p = 0; p < nparts; p++) acc += __ldcg(F.T2.partials + (size_t)p * m + i);
                const double mine = (F.Ws ? __dmul_rn(x[i], F.Ws[i]) : 0.0) + acc;
                for (int r = 0; r < nranks; r++) {
                    ulonglong2* dst = reinterpret_cast<ulonglong2*>(
                                          reinterpret_cast<char*>(F.peers[r]) + F.xll_off) +
                                      par + (size_t)F.rank * F.xmpad + i;
                    xll_store(dst, gen, mine);
                }
            }
            const ulonglong2* rec = reinterpret_cast<const ulonglong2*>(
                                        reinterpret_cast<const char*>(F.peers[F.rank]) + F.xll_off) + par;
            for (int i = i0 + tid; i < i1; i += nthr) {
                double yv = 0.0;
                for (int rb = 0; rb < nranks; rb += 8) {
                    double part[8];
                    bool ok[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        part[u] = 0.0;
                        ok[u] = rb + u >= nranks ||
                                xll_load(rec + (size_t)(rb + u) * F.xmpad + i, gen, &part[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        long long spins = 0;
                        while (!ok[u]) {
                            ok[u] = xll_load(rec + (size_t)(rb + u) * F.xmpad + i, gen, &part[u]);
                            if ((++spins & 0xffff) == 0) {
                                // a peer that never arrives must not hang this GPU
                                if (spins > (1ll << 24) || __ldcg(F.abort_word) != 0.0) {
                                    *F.abort_word = 2.0;
                                    break;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (rb + u < nranks) yv += part[u];
                }
                lhs[i] = yv;
                dot += __dmul_rn(x[i], yv);
            }
        } else if (nranks > 1) {
            // Column shards, pull exchange (IPXGPU_XCHG=pull): this rank's partial product of the slice goes to its
            // exchange buffer; the same CTA of every rank then sums the ranks'
            // partials of the slice in rank order (bit-identical on all ranks)
            // with P2P loads. Synchronisation is per slice: a flag per (rank, CTA)
            // in every peer's buffer, no collective and no extra grid barrier.
            const unsigned gen = F.xgen_base + (unsigned)s_applies;
            const size_t off = (size_t)(gen & 1u) * F.xmpad;
            double* mine = F.peers[F.rank] + off;
            for (int i = i0 + tid; i < i1; i += nthr) {
                double acc = 0.0;
#pragma unroll 8
                for (int p = 0; p < nparts; p++) acc += __ldcg(F.T2.partials + (size_t)p * m + i);
                mine[i] = (F.Ws ? __dmul_rn(x[i], F.Ws[i]) : 0.0) + acc;
            }
            __syncthreads();
            const size_t flag_off = 2 * F.xmpad;  // in doubles
            if (tid == 0) __threadfence_system();
            __syncthreads();
            if (tid < nranks) {
                // signal rank `tid` that slice blockIdx.x of this rank is ready ...
                unsigned* remote = reinterpret_cast<unsigned*>(F.peers[tid] + flag_off) +
                                   (size_t)F.rank * gridDim.x + blockIdx.x;
                asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(gen) : "memory");
                // ... and wait for rank `tid`'s slice
                const unsigned* local = reinterpret_cast<const unsigned*>(F.peers[F.rank] + flag_off) +
                                        (size_t)tid * gridDim.x + blockIdx.x;
                unsigned now;
                long long spins = 0;
                do {
                    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(now) : "l"(local) : "memory");
                    if ((++spins & 0xffff) == 0) {
                        // a peer that never arrives must not hang this GPU
                        if (spins > (1ll << 24) || __ldcg(F.abort_word) != 0.0) {
                            *F.abort_word = 2.0;
                            break;
                        }
                    }
                } while ((int)(now - gen) < 0);
                __threadfence_system();
            }
            __syncthreads();
            for (int i = i0 + tid; i < i1; i += nthr) {
                // All peers' loads go out before the first sum waits (one NVLink round
                // trip per group of 8 ranks instead of one per rank); summed in rank order.
                double yv = 0.0;
                for (int rb = 0; rb < nranks; rb += 8) {
                    double part[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        part[u] = 0.0;
                        if (rb + u < nranks)
                            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];"
                                         : "=d"(part[u])
                                         : "l"(F.peers[rb + u] + off + i));
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (rb + u < nranks) yv += part[u];
                }
                lhs[i] = yv;
                dot += __dmul_rn(x[i], yv);
            }
        }
        sync_stage(kStCombine, dot, 0.0, extra_max);
        stamp_lead(kSlotOp);
        return r0;
    };

    // ---- initialisation (reference :33-40, :118-127) ----
    if (!F.zero_start) {
        apply(v.y, v.Cs, false, 0.0);  // own slice of C*y, read back by the same threads below
    }
    {
        double rs = 0.0, mx = 0.0;
        for (int i = i0 + tid; i < i1; i += nthr) {
            const double ri = F.zero_start ? F.rhs[i] : F.rhs[i] - __ldcg(v.Cs + i);
            v.r[i] = ri;
            v.p[i] = 0.0;
            v.Cp[i] = 0.0;
            if (precond) {
                const double si = ri / v.diag[i];
                v.s[i] = si;
                rs += __dmul_rn(si, ri);
            }
            const double sc = v.resscale ? __dmul_rn(v.resscale[i], ri) : ri;
            mx = fmax(mx, fabs(sc));
        }
        sync_stage(kStInit, rs, 0.0, mx);
        if (tid == 0) {
            s_st.rsdot_prev = r0;
            s_st.resnorm = r2;
        }
        stamp_lead(precond ? kSlotPre : kSlotVec);
        __syncthreads();
    }
    const double* sv = precond ? v.s : v.r;
    {
        const double dot = apply(sv, v.Cs, false, 0.0);
        if (tid == 0) {
            s_st.cdot = dot;
            s_st.beta = 0.0;
        }
        __syncthreads();
    }

    for (;;) {
        // ---- direction (reference :79-80 / :181-207) and the tests at the top
        //      of the next pass (:50-71 / :137-172) ----
        const double beta = s_st.beta;
        const long long iter = s_st.iter;
        const bool recompute = precond && iter > 0 && (iter % 5 == 0);
        double pd = 0.0, rs = 0.0;
        for (int i = i0 + tid; i < i1; i += nthr) {
            const double pn = sv[i] + __dmul_rn(beta, v.p[i]);
            const double cpn = __ldcg(v.Cs + i) + __dmul_rn(beta, v.Cp[i]);
            v.p[i] = pn;
            v.Cp[i] = cpn;
            if (precond) {
                const double d = v.diag[i];
                const double qi = cpn / d;
                v.q[i] = qi;
                pd += __dmul_rn(qi, cpn);
                if (recompute) {
                    const double ri = v.r[i];
                    const double sn = ri / d;
                    v.s[i] = sn;
                    rs += __dmul_rn(sn, ri);
                }
            } else {
                pd += __dmul_rn(cpn, cpn);
            }
        }
        if (lead && (iter & 3) == 0 && F.abort_flag && *F.abort_flag) *F.abort_word = 1.0;
        sync_stage(kStDirection, pd, rs, 0.0);
        const double tpd = r0, trs = r1;
        const bool aborted = __ldcg(F.abort_word) != 0.0;
        if (tid == 0) {
            int done = 0, err = 0;
            if (recompute) {
                if (trs >= s_st.rsdot_prev) {
                    err = 204;
                    done = 1;
                } else {
                    s_st.rsdot_prev = trs;
                }
            }
            if (!done) {
                const double resnorm = s_st.resnorm;
                if (lead && s_st.hist && iter < s_st.hist_cap) s_st.hist[iter] = resnorm;
                s_st.pdot = tpd;
                if (resnorm <= s_st.tol) {
                    done = 1;
                } else if (iter == s_st.maxiter) {
                    err = 201;
                    done = 1;
                } else if (s_st.cdot <= 0.0) {
                    err = 202;
                    done = 1;
                } else if (precond && tpd <= 0.0) {
                    err = 203;
                    done = 1;
                } else {
                    const double alpha = s_st.cdot / tpd;
                    if (!isfinite(alpha)) {
                        err = 205;
                        done = 1;
                    }
                    s_st.alpha = alpha;
                }
            }
            if (aborted) done = 1;
            s_st.errflag = err;
            s_st.done = done;
            if (lead) stamp(&s_st, precond ? kSlotPre : kSlotVec);
        }
        __syncthreads();
        if (s_st.done) break;

        // ---- update (reference :72-73 / :173-175) ----
        double upd_max = 0.0;
        {
            const double alpha = s_st.alpha;
            double mx = 0.0;
            for (int i = i0 + tid; i < i1; i += nthr) {
                v.y[i] = v.y[i] + __dmul_rn(alpha, v.p[i]);
                const double ri = v.r[i] - __dmul_rn(alpha, v.Cp[i]);
                v.r[i] = ri;
                if (precond) v.s[i] = v.s[i] - __dmul_rn(alpha, v.q[i]);
                const double sc = v.resscale ? __dmul_rn(v.resscale[i], ri) : ri;
                mx = fmax(mx, fabs(sc));
            }
            ++x_gen;
            if (flags_on) {
                // no grid barrier: publish this CTA's slice, the residual norm rides the
                // reduction at the end of the apply
                __syncthreads();
                if (tid == 0) {
                    fence_proxy_async();
                    __threadfence();
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(F.x_ready + blockIdx.x),
                                 "r"(x_gen)
                                 : "memory");
                }
                upd_max = mx;
            } else {
                sync_stage(kStUpdate, 0.0, 0.0, mx);
                if (tid == 0) s_st.resnorm = r2;
            }
            stamp_lead(kSlotVec);
        }
        // ---- C.Apply and its scalar step (reference :75-81 / :176-184) ----
        {
            const double dot = apply(sv, v.Cs, true, flags_on ? upd_max : 0.0);
            if (flags_on && tid == 0) s_st.resnorm = r2;
            if (tid == 0) {
                s_st.beta = dot / s_st.cdot;
                s_st.cdot = dot;
                s_st.iter += 1;
            }
            __syncthreads();
        }
    }

    if (lead) {
        s_st.applies = s_applies;
        *F.st = s_st;
        publish(F.st);
    }
}


// The record exchange as a kernel of its own, for sharded contexts that run the
// launch-per-stage CR loop (generic sweeps, e.g. config 5 with m = 10^6): y (this rank's
// partial product, slack term included on rank 0) becomes the sum over the ranks, in rank
// order, on every rank; y[m] = x'y; then the scalar step that follows the apply. Replaces
// ncclAllReduce(m+1) + cr_after_apply_kernel. Same two shapes and the same records as the
// combine stage of pcr_fused_kernel. Grid-stride; the grid must be co-resident on every rank
// (<= 8 CTAs per SM), because a thread waits for the peers' threads of the same element.
struct XchgArgs {
    double* y;        // m+1
    const double* x;  // m
    int m;
    int nranks, rank;
    double* const* peers;
    size_t xmpad, xll_off;
    unsigned gen;
    int two_phase;
    int mode, slot;   // after_apply
    double* abort_word;
};

__global__ void __launch_bounds__(kBlock)
xchg_records_kernel(XchgArgs A, Reduce red, CrState* st) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    if (st != nullptr && st->done) return;
    const unsigned gen = A.gen;
    const size_t par = (size_t)(gen & 1u) * (size_t)A.nranks * A.xmpad;
    const size_t fin_off = A.xll_off + 2 * (size_t)A.nranks * A.xmpad * 16 +
                           (size_t)(gen & 1u) * A.xmpad * 16;
    auto wait_rec = [&](const ulonglong2* p, double* val) {
        long long spins = 0;
        while (!xll_load(p, gen, val)) {
            if ((++spins & 0xffff) == 0) {
                if (spins > (1ll << 24) || __ldcg(A.abort_word) != 0.0) {
                    *A.abort_word = 2.0;  // a peer that never arrives must not hang this GPU
                    break;
                }
            }
        }
    };
    const ulonglong2* rec = reinterpret_cast<const ulonglong2*>(
                                reinterpret_cast<const char*>(A.peers[A.rank]) + A.xll_off) + par;
    const ulonglong2* fin = reinterpret_cast<const ulonglong2*>(
        reinterpret_cast<const char*>(A.peers[A.rank]) + fin_off);
    const int per = (A.m + A.nranks - 1) / A.nranks;  // two-phase: rows [r*per, (r+1)*per) -> rank r
    const int first = blockIdx.x * kBlock + threadIdx.x, stride = gridDim.x * kBlock;
    double dot = 0.0;
    // Three passes over this thread's elements, so that its pushes, its sums and its final
    // reads are each in flight together (one NVLink hop per pass, not per element).
    if (A.two_phase) {
        for (int i = first; i < A.m; i += stride)
            xll_store(reinterpret_cast<ulonglong2*>(reinterpret_cast<char*>(A.peers[i / per]) +
                                                    A.xll_off) + par + (size_t)A.rank * A.xmpad + i,
                      gen, A.y[i]);
        for (int i = first; i < A.m; i += stride) {
            if (i / per != A.rank) continue;
            double tot = 0.0;
            for (int r = 0; r < A.nranks; r++) {
                double part;
                wait_rec(rec + (size_t)r * A.xmpad + i, &part);
                tot += part;
            }
            for (int r = 0; r < A.nranks; r++)
                xll_store(reinterpret_cast<ulonglong2*>(reinterpret_cast<char*>(A.peers[r]) + fin_off) + i,
                          gen, tot);
        }
        for (int i = first; i < A.m; i += stride) {
            double yv;
            wait_rec(fin + i, &yv);
            A.y[i] = yv;
            dot += __dmul_rn(A.x[i], yv);
        }
    } else {
        for (int i = first; i < A.m; i += stride) {
            const double mine = A.y[i];
            for (int r = 0; r < A.nranks; r++)
                xll_store(reinterpret_cast<ulonglong2*>(reinterpret_cast<char*>(A.peers[r]) +
                                                        A.xll_off) + par + (size_t)A.rank * A.xmpad + i,
                          gen, mine);
        }
        for (int i = first; i < A.m; i += stride) {
            double yv = 0.0;
            for (int r = 0; r < A.nranks; r++) {
                double part;
                wait_rec(rec + (size_t)r * A.xmpad + i, &part);
                yv += part;
            }
            A.y[i] = yv;
            dot += __dmul_rn(A.x[i], yv);
        }
    }
    const double mine_dot = block_sum(dot, s_red);
    double ts, ts2, tm;
    if (grid_reduce(red, mine_dot, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) && threadIdx.x == 0) {
        A.y[A.m] = ts;
        if (st && !(A.mode == kApplyPlain && A.slot == kSlotNone)) after_apply(st, A.mode, ts, A.slot);
    }
}

}  // namespace ipxgpu
