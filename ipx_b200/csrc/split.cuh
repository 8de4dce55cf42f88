// Level-scheduled sparse triangular solves and the basis-preconditioned
// operator C = I + inv(B) N N' inv(B')  (reference src/sparse_matrix.cc:224-311,
// src/splitted_normal_matrix.cc:18-117).
//
// The reference solves L and U column-wise (scatter) and U', L' row-wise
// (gather), all strictly sequentially. Here all four solves are in gather
// form: every row i reads already-solved entries x[k], so the rows of one
// dependency level are independent. Rows are bucketed by level on the host.
// One persistent cooperative kernel per solve (tri_syncfree_kernel), no level barriers: the
// rows are sorted by level and dealt to the grid's warps in that order; a warp owns a row,
// loads 32 of its entries at a time (coalesced index/value loads), every lane waits for the
// ready flag of the row its entry depends on, gathers x, and the 32 rounded products are
// folded into the row's value one after the other with shuffles, i.e. in exactly the order
// the reference applies them (rounded product, then add/subtract), so the solves are
// bit-identical to the CPU loops. The finished row publishes x[i] and its flag (release).
// Dependencies sit at earlier positions of the order and every warp walks its positions in
// increasing order, so the earliest unfinished row can always proceed: no deadlock as long
// as the grid is co-resident (cooperative launch). A dependency chain of length d costs d
// flag round trips through L2 (~1 us each), not d kernel launches or grid barriers; rows
// that merely follow each other in the order overlap.
// (The first version gave a whole run of narrow levels to one CTA with one THREAD per row:
// 1 s per apply on a 50,000-row basis whose factors hold 11 M entries and a dense trailing
// block of 2000 rows - 25x slower than the CPU loops.)
#pragma once

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "context.cuh"
#include "spmv_kernels.cuh"

namespace ipxgpu {

constexpr int kTriBlock = 1024;      // CTA of the merged-level kernel
constexpr int kTriWideRows = 4096;   // levels at least this wide get a grid
constexpr int kTriWarps = 12;        // warps per CTA of the sync-free kernel (16 KB of stash each)

struct TriStep {
    int wide;        // 1: one level on a full grid; 0: levels [l0, l1) in one CTA
    int l0, l1;
    int r0, r1;      // positions in `order`
};

// One triangular system in gather form (device view, passed by value).
struct TriDev {
    int dim = 0;
    int subtract_seq = 0;  // 1: v -= a*x per entry (L, U column sweeps of the
                           // reference); 0: v = x[i] - sum (U', L' row sweeps)
    int reverse = 0;       // a row's entries resolve back to front (L': the reference sums a
                           // column of L from the diagonal downwards, the solve runs upwards).
                           // 1: summed front to back all the same (the reference's order, a
                           // serial chain of additions AFTER the last dependency arrived);
                           // 2: summed back to front, in the order the dependencies resolve
    int prefix = 1;        // fold what precedes a chunk's first missing entry before waiting
    int lane_sums = 1;     // rows longer than kTriStash entries: all pieces but the one whose
                           // dependencies resolve last are summed per lane (32 partial sums,
                           // folded in lane order) instead of entry by entry; 0: reference order
    int mailbox = 1;       // rows solved by a warp of the same CTA are read from shared memory
    unsigned pause_cap = 32;   // longest pause (ns) between two rounds of looks at global records
    unsigned long long* trace = nullptr;  // tuning only: [2*dim] globaltimer at row start / finish
    int* pos = nullptr;    // [dim] position of row i in `order` (inverse permutation)
    int* ptr = nullptr;    // [dim+1] off-diagonal entries of row i
    int* idx = nullptr;
    double* val = nullptr;
    double* diag = nullptr;  // nullptr: unit diagonal
    int* order = nullptr;    // rows sorted by level
    int* level_ptr = nullptr;
};

struct TriSystem {
    TriDev d;
    int nlevels = 0;
    std::vector<TriStep> steps;
};

struct SplitOperator {
    int dim = 0;
    TriSystem sys[4];  // 0: L, 1: U, 2: U', 3: L'
    bool lu_loaded = false;
    bool prepared = false;
    double* W2 = nullptr;      // nloc + m squared nonbasic scales (0 elsewhere)
    int* rinv = nullptr;       // m: row i of AI -> position in the permuted system
    unsigned char* free_mask = nullptr;  // m
    double* work = nullptr;    // m
    double* xun = nullptr;     // m
    double* yun = nullptr;     // m+1
    // KKTSolverBasis::_Solve on the device (ipxgpu_kktbasis_prepare)
    bool kkt_ready = false;
    int num_free = 0;
    int* basic_var = nullptr;     // m: variable j = basis[colperm[k]] at pivot position k
    int* colperm = nullptr;       // m: basis position p = colperm[k]
    double* basic_scale = nullptr;  // m: colscale[j] of BASIC positions, 1 for BASIC_FREE
    double* tk = nullptr;         // m, pivot-position space
    double* wrow = nullptr;       // m, row space
};

__device__ __forceinline__ void tri_row(const TriDev& T, int i, double* x) {
    const int b = T.ptr[i], e = T.ptr[i + 1];
    double v = x[i];
    if (T.subtract_seq) {
        for (int p = b; p < e; p++) v = v - __dmul_rn(T.val[p], x[T.idx[p]]);
    } else {
        double d = 0.0;
        for (int p = b; p < e; p++) d = d + __dmul_rn(x[T.idx[p]], T.val[p]);
        v = v - d;
    }
    if (T.diag) v = v / T.diag[i];
    x[i] = v;
}

// The reference's column sweeps divide BEFORE scattering (x[j] /= U_jj, then
// x[i] -= U_ij x[j]), i.e. row i ends as (x_i - sum)/U_ii as well; the order of
// subtractions is what tri_row reproduces.

__global__ void __launch_bounds__(kBlock)
tri_wide_kernel(TriDev T, int r0, int r1, double* x, const CrState* st) {
    if (st && st->done) return;
    const int r = r0 + blockIdx.x * kBlock + threadIdx.x;
    if (r < r1) tri_row(T, T.order[r], x);
}

__global__ void __launch_bounds__(kTriBlock)
tri_levels_kernel(TriDev T, int l0, int l1, double* x, const CrState* st) {
    if (st && st->done) return;
    for (int l = l0; l < l1; l++) {
        const int rb = T.level_ptr[l], re = T.level_ptr[l + 1];
        for (int r = rb + threadIdx.x; r < re; r += kTriBlock) tri_row(T, T.order[r], x);
        __syncthreads();
    }
}

// Solved entries travel as 16-byte records of two self-validating 64-bit words
// {generation : 32 | low half of x}, {generation : 32 | high half of x} (the layout of NCCL's LL
// protocol): a 64-bit word is read and written atomically, so a reader that finds this solve's
// generation in both words holds the final x[j] - no separate flag, no fence on either side,
// one L2 round trip per dependency instead of two.
__device__ __forceinline__ bool tri_ll_load(const ulonglong2* p, unsigned gen, double* val) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];"
                 : "=l"(w0), "=l"(w1)
                 : "l"(p)
                 : "memory");
    *val = __longlong_as_double((long long)(((w1 & 0xffffffffull) << 32) | (w0 & 0xffffffffull)));
    return (unsigned)(w0 >> 32) == gen && (unsigned)(w1 >> 32) == gen;
}
__device__ __forceinline__ void tri_ll_store(ulonglong2* p, unsigned gen, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long w0 = ((unsigned long long)gen << 32) | (bits & 0xffffffffull);
    const unsigned long long w1 = ((unsigned long long)gen << 32) | (bits >> 32);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1)
                 : "memory");
}

// One row by one warp; the same arithmetic, in the same order, as tri_row. x is read and
// written through L2 (other SMs solved the rows it depends on); flags[j] == gen says that
// x[j] of this solve is final.
//
// Two phases per row (per piece of kTriStash entries for longer rows):
//  A. never blocks: in batches of kTriBatch chunks of 32 entries the index/value loads go out
//     together, then the flag loads, then (after one fence) the x loads of every entry whose
//     flag was up - three memory round trips per 256 entries. Each lane parks its rounded
//     products in its own shared-memory slots and remembers which of its entries were not
//     ready.
//  B. in order: chunk by chunk the 32 products are folded into the row value with shuffles.
//     A chunk with entries that were not ready waits for them first: all lanes look once, then
//     only the first lane still waiting polls (with a growing pause), so thousands of waiting
//     warps do not saturate the L2 slice that holds the flags.
// So whatever a row can prepare is prepared while it waits for its last dependency, wherever
// that sits in the summation order: in the L' solve the reference sums a row starting with
// the entry next to the diagonal, i.e. the dependency that resolves LAST, and without phase A
// a 1000-entry row of a dense trailing block paid all its memory round trips after that.
constexpr int kTriBatch = 8;
constexpr int kTriStash = 2048;  // parked products per warp (16 KB)
constexpr int kTriMailDepth = 8; // records per warp in the CTA's shared-memory mailbox

// Mailbox of a CTA: the last kTriMailDepth rows each of its warps solved, as self-validating
// records {round+1 : 32 | half of x} x 2 in shared memory. Row positions are dealt round robin
// over the grid's warps, so a dependency chain (one row per level) runs through the 12 warps of
// a CTA before it moves to the next CTA: 11 of 12 links can be handed over through shared memory
// (tens of cycles) instead of an L2 round trip. The global record is always written as well; a
// reader that finds its slot overwritten (the producer is kTriMailDepth rounds ahead) takes
// that one.
__device__ __forceinline__ void tri_mail_store(ulonglong2* slot, unsigned tag, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    ulonglong2 rec;
    rec.x = ((unsigned long long)tag << 32) | (bits & 0xffffffffull);
    rec.y = ((unsigned long long)tag << 32) | (bits >> 32);
    *reinterpret_cast<volatile unsigned long long*>(&slot->x) = rec.x;
    *reinterpret_cast<volatile unsigned long long*>(&slot->y) = rec.y;
}
// 1: value read, 0: not there yet, -1: overwritten by a later round (take the global record)
__device__ __forceinline__ int tri_mail_load(const ulonglong2* slot, unsigned tag, double* val) {
    const unsigned long long w0 = *reinterpret_cast<const volatile unsigned long long*>(&slot->x);
    const unsigned long long w1 = *reinterpret_cast<const volatile unsigned long long*>(&slot->y);
    const unsigned t0 = (unsigned)(w0 >> 32), t1 = (unsigned)(w1 >> 32);
    if (t0 == tag && t1 == tag) {
        *val = __longlong_as_double((long long)(((w1 & 0xffffffffull) << 32) | (w0 & 0xffffffffull)));
        return 1;
    }
    return (t0 > tag || t1 > tag) ? -1 : 0;
}

// Folds lanes [lo, hi) of `prod` into the running value, lane after lane (sub: v -= p, else
// v += p); down: from hi-1 to lo. The bounds are uniform over the warp. The shuffles of a group go
// out back to back and only the chain of additions stays serial (a rolled loop pays shuffle
// latency + add per entry). Lanes beyond hi contribute -0.0, the one addend that leaves every
// value - both zeros included - as it is: no select sits on the chain of additions, and the few
// entries that remain after a row's last dependency arrived cost a few additions, not 32.
__device__ __forceinline__ double tri_fold(double v, double prod, int lo, int hi, bool sub,
                                           bool down) {
    const double sp = sub ? -prod : prod;  // v - p == v + (-p) exactly
    if (lo == 0 && hi == 32) {
#pragma unroll
        for (int k = 0; k < 32; k++) v = v + __shfl_sync(0xffffffffu, sp, down ? 31 - k : k);
        return v;
    }
    for (int k0 = lo; k0 < hi; k0 += 8) {
        double p[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int kk = k0 + u;
            const int src = down ? hi - 1 - (kk - lo) : kk;
            p[u] = __shfl_sync(0xffffffffu, sp, src & 31);
            if (kk >= hi) p[u] = -0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) v = v + p[u];
    }
    return v;
}

__device__ __forceinline__ void tri_row_warp(const TriDev& T, int i, double* x, ulonglong2* ll,
                                             unsigned gen, int lane, double* stash, unsigned* err,
                                             ulonglong2* mail, int nwarps, int warp, int round) {
    const int b = T.ptr[i], e = T.ptr[i + 1];
    const bool sub = T.subtract_seq != 0;
    const bool down = T.reverse == 2;  // pieces, chunks and lanes from the back
    if (T.trace != nullptr && lane == 0) T.trace[2 * i] = globaltimer();
    double v = __ldcg(x + i);
    double d = 0.0;
    const int npieces = (e - b + kTriStash - 1) / kTriStash;
    const bool lanesum = T.lane_sums != 0 && T.reverse != 1 && npieces > 1;
    double part = 0.0;  // lanesum: this lane's share of the pieces before the last one
    for (int pc = 0; pc < npieces; pc++) {
        const int q0 = b + (down ? npieces - 1 - pc : pc) * kTriStash;
        const int qe = min(e, q0 + kTriStash);
        unsigned long long mypend = 0ull;  // bit c: my entry of chunk c was not ready in phase A
        for (int p0 = q0; p0 < qe; p0 += 32 * kTriBatch) {
            int j[kTriBatch];
            double a[kTriBatch], xv[kTriBatch];
            bool ready[kTriBatch];
#pragma unroll
            for (int u = 0; u < kTriBatch; u++) {
                const int p = p0 + 32 * u + lane;
                const bool active = p < qe;
                j[u] = active ? __ldg(T.idx + p) : -1;
                a[u] = active ? __ldg(T.val + p) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kTriBatch; u++) {
                xv[u] = 0.0;
                ready[u] = j[u] < 0 || tri_ll_load(ll + j[u], gen, &xv[u]);
            }
            const int c0 = (p0 - q0) >> 5;
#pragma unroll
            for (int u = 0; u < kTriBatch; u++) {
                if (p0 + 32 * u >= qe) break;  // uniform
                stash[(c0 + u) * 32 + lane] = (j[u] >= 0 && ready[u]) ? __dmul_rn(a[u], xv[u]) : 0.0;
                if (!ready[u]) mypend |= 1ull << (c0 + u);
            }
        }
        const int nchunks = (qe - q0 + 31) >> 5;
        // Waits for the entries of chunk c that were not ready in phase A and returns this
        // lane's product of the chunk.
        auto resolve = [&](int c, double prod) -> double {
            const bool mine = (mypend >> c) & 1ull;
            if (__ballot_sync(0xffffffffu, mine) == 0u) return prod;
            int jj = 0;
            double aa = 0.0;
            const ulonglong2* slot = nullptr;  // mailbox slot of a row solved in this CTA
            unsigned want = 0;
            if (mine) {
                jj = __ldg(T.idx + q0 + 32 * c + lane);
                aa = __ldg(T.val + q0 + 32 * c + lane);
                if (mail != nullptr) {
                    // position p of the order is solved by warp p % nwarps in its round p / nwarps
                    const int pj = __ldg(T.pos + jj);
                    const int g = pj % nwarps, rj = pj / nwarps;
                    if (g / kTriWarps == (int)blockIdx.x) {
                        slot = mail + (g % kTriWarps) * kTriMailDepth + (rj % kTriMailDepth);
                        want = (unsigned)rj + 1u;
                    }
                }
            }
            bool ok = !mine;
            double xj = 0.0;
            // one look: the mailbox where the row is this CTA's, the global record otherwise
            auto look = [&]() -> bool {
                if (slot != nullptr) {
                    const int r = tri_mail_load(slot, want, &xj);
                    if (r > 0) return true;
                    if (r == 0) return false;
                    slot = nullptr;  // overwritten: the global record is there
                }
                return tri_ll_load(ll + jj, gen, &xj);
            };
            // All lanes that miss an entry look together, then round after round the first few
            // of those still missing: entries arrive in lane order at the head of a dependency
            // chain, and in batches where a level is wide. An entry is seen within one L2 round
            // trip of its arrival, entries that arrived together cost one round trip, and a
            // waiting warp keeps at most `window` looks in flight - hundreds of warps wait for
            // the same few records at the head of a chain. (Round 1 let only the first missing
            // lane poll, after a look by all: two round trips per hop, tools/tri_trace.py.)
            unsigned pause = 0, spins = 0;
            int window = 4;
            if (!ok) ok = look();
            unsigned pending = __ballot_sync(0xffffffffu, !ok);
            while (pending != 0u) {
                // the window sits where the next arrivals are: at the lowest missing lanes, or
                // at the highest ones where the row is summed from the back
                const int first = __ffs(pending) - 1, last = 31 - __clz(pending);
                const bool in_window = T.reverse != 0 ? lane > last - window : lane < first + window;
                if (!ok && in_window) ok = look();
                const unsigned now = __ballot_sync(0xffffffffu, !ok);
                // the whole window arrived at once: a batch, look at twice as many next time
                const unsigned win_mask = __ballot_sync(0xffffffffu, in_window);
                if ((now & win_mask) == 0u) {
                    window = min(2 * window, 32);
                    pause = 0;
                } else if (now == pending) {
                    if (pause) __nanosleep(pause);
                    pause = min(2 * pause + 16u, T.pause_cap);
                    window = 4;
                }
                pending = now;
                // A dependency that never resolves (cannot happen with a valid level order) must
                // not hang the GPU: give up after ~2 s and flag the solve.
                if (++spins > (1u << 23)) {
                    if (lane == 0) *err = 1u;
                    break;
                }
            }
            if (mine) prod = __dmul_rn(aa, xj);
            return prod;
        };
        if (lanesum && pc < npieces - 1) {
            // A serial chain of fp64 additions advances by one entry every ~14 cycles: a
            // linking row of a block-angular basis with 40,000 entries holds everything that
            // depends on it for 300 us (the CPU needs 50). Here every lane sums its own entries
            // of the piece, 64 additions deep.
            for (int c = 0; c < nchunks; c++) {
                double prod = stash[c * 32 + lane];
                if (__ballot_sync(0xffffffffu, (mypend >> c) & 1ull) != 0u) prod = resolve(c, prod);
                part = part + prod;
            }
            continue;
        }
        if (lanesum) {
            // the last piece (in the order of the sum): first the 32 partial sums, lane by lane
            if (sub) v = tri_fold(v, part, 0, 32, true, false);
            else d = tri_fold(d, part, 0, 32, false, false);
        }
        if (T.reverse == 1) {
            // The dependencies of the LAST chunk resolve first (see above): wait for the
            // chunks back to front and park the products, then sum front to back.
            for (int c = nchunks - 1; c >= 0; c--) {
                if (__ballot_sync(0xffffffffu, (mypend >> c) & 1ull) == 0u) continue;
                stash[c * 32 + lane] = resolve(c, stash[c * 32 + lane]);
            }
            mypend = 0ull;
        }
        for (int cc = 0; cc < nchunks; cc++) {
            const int c = down ? nchunks - 1 - cc : cc;
            const int cnt = min(32, qe - (q0 + 32 * c));
            double prod = stash[c * 32 + lane];
            double acc = sub ? v : d;
            const unsigned missing = __ballot_sync(0xffffffffu, (mypend >> c) & 1ull);
            if (missing == 0u) {
                acc = tri_fold(acc, prod, 0, cnt, sub, down);
            } else if (!T.prefix) {
                prod = resolve(c, prod);
                acc = tri_fold(acc, prod, 0, cnt, sub, down);
            } else if (!down) {
                // what precedes the first missing entry is folded while it is awaited
                const int first = __ffs(missing) - 1;
                acc = tri_fold(acc, prod, 0, first, sub, false);
                prod = resolve(c, prod);
                acc = tri_fold(acc, prod, first, cnt, sub, false);
            } else {
                const int last = 31 - __clz(missing);
                acc = tri_fold(acc, prod, last + 1, cnt, sub, true);
                prod = resolve(c, prod);
                acc = tri_fold(acc, prod, 0, last + 1, sub, true);
            }
            if (sub) v = acc; else d = acc;
        }
    }
    if (!sub) v = v - d;
    if (T.diag) v = v / __ldg(T.diag + i);
    if (lane == 0) {
        tri_ll_store(ll + i, gen, v);  // what the dependent rows read
        if (mail != nullptr)
            tri_mail_store(mail + warp * kTriMailDepth + (round % kTriMailDepth),
                           (unsigned)round + 1u, v);
        x[i] = v;                      // the result
        if (T.trace != nullptr) T.trace[2 * i + 1] = globaltimer();
    }
}

// Whole solve in one cooperative launch (the grid must be co-resident, see above).
__global__ void __launch_bounds__(kTriWarps * 32, 1)
tri_syncfree_kernel(TriDev T, double* x, ulonglong2* ll, unsigned gen, unsigned* err,
                    const CrState* st) {
    extern __shared__ __align__(16) double tri_stash[];  // kTriWarps * kTriStash, then the mailbox
    if (st && st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * kTriWarps + warp;
    const int nwarps = gridDim.x * kTriWarps;
    double* stash = tri_stash + (size_t)warp * kTriStash;
    ulonglong2* mail = nullptr;
    if (T.mailbox) {
        mail = reinterpret_cast<ulonglong2*>(tri_stash + (size_t)kTriWarps * kTriStash);
        for (int k = threadIdx.x; k < kTriWarps * kTriMailDepth; k += kTriWarps * 32)
            mail[k] = make_ulonglong2(0ull, 0ull);
        __syncthreads();
    }
    int round = 0;
    for (int r = gwarp; r < T.dim; r += nwarps, round++)
        tri_row_warp(T, T.order[r], x, ll, gen, lane, stash, err, mail, nwarps, warp, round);
}

__global__ void __launch_bounds__(kBlock)
gather_perm_kernel(int m, const int* __restrict__ perm, const double* __restrict__ src,
                   double* __restrict__ dst, CrState* st, int slot) {
    if (st && st->done) return;
    if (st && blockIdx.x == 0 && threadIdx.x == 0) stamp(st, slot);
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < m; i += gridDim.x * kBlock)
        dst[i] = src[perm[i]];
}

__global__ void __launch_bounds__(kBlock)
scatter_perm_kernel(int m, const int* __restrict__ perm, const double* __restrict__ src,
                    double* __restrict__ dst, CrState* st, int slot) {
    if (st && st->done) return;
    if (st && blockIdx.x == 0 && threadIdx.x == 0) stamp(st, slot);
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < m; i += gridDim.x * kBlock)
        dst[perm[i]] = src[i];
}

// lhs += rhs; lhs[free] = 0; lhs[m] = rhs'lhs
// (reference src/splitted_normal_matrix.cc:112-116).
__global__ void __launch_bounds__(kBlock)
split_finish_kernel(int m, const double* __restrict__ rhs, double* __restrict__ lhs,
                    const unsigned char* __restrict__ free_mask, Reduce red, int mode,
                    CrState* st) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    if (st && st->done) return;
    double acc = 0.0;
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < m; i += gridDim.x * kBlock) {
        double v = lhs[i] + rhs[i];
        if (free_mask[i]) v = 0.0;
        lhs[i] = v;
        acc += __dmul_rn(rhs[i], v);
    }
    const double b = block_sum(acc, s_red);
    double ts, ts2, tm;
    if (grid_reduce(red, b, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) && threadIdx.x == 0) {
        lhs[m] = ts;
        if (st) after_apply(st, mode, ts, kSlotB);
    }
}

__global__ void __launch_bounds__(kBlock)
square_kernel(long long n, const double* __restrict__ a, double* __restrict__ out) {
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < n;
         j += (long long)gridDim.x * kBlock)
        out[j] = __dmul_rn(a[j], a[j]);
}

// ---- elementwise steps of KKTSolverBasis::_Solve (reference src/kkt_solver_basis.cc:75-194) ----
// k = pivot position, j = basic_var[k] the variable there, d = basic_scale[k]. The resident U
// carries the column scales of the BASIC variables (SplittedNormalMatrix::Prepare,
// src/splitted_normal_matrix.cc:30-39), so with Us = U D:  inverse(B) v = D Us^{-1} L^{-1} P v and
// inverse(B') v = P' L^{-T} Us^{-T} D v; the divisions / multiplications by d of :128-137 and
// :166-177 cancel against that D and are not performed.

// out[k] = free[k] ? a[basic_var[k]] : (src ? src[k] : 0)   (:88-96 with src == nullptr, :165-177)
__global__ void __launch_bounds__(kBlock)
kb_slot_free_kernel(int m, const int* __restrict__ basic_var, const unsigned char* __restrict__ free_mask,
                    const double* __restrict__ a, const double* __restrict__ src,
                    double* __restrict__ out) {
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < m; k += gridDim.x * kBlock)
        out[k] = free_mask[k] ? a[basic_var[k]] : (src ? src[k] : 0.0);
}

// Slack part of the masked column sweeps: us[i] = nb2s[i] * (as[i] - (w ? w[i] : 0)); then
// init[i] = sign > 0 ? us[i] : b[i] - us[i]  (:100-119, :180-189 for the identity columns).
__global__ void __launch_bounds__(kBlock)
kb_slack_kernel(int m, const double* __restrict__ nb2s, const double* __restrict__ as,
                const double* __restrict__ w, const double* __restrict__ b, double sign,
                double* __restrict__ us, double* __restrict__ init) {
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < m; i += gridDim.x * kBlock) {
        const double u = __dmul_rn(nb2s[i], as[i] - (w ? w[i] : 0.0));
        us[i] = u;
        init[i] = sign > 0 ? u : b[i] - u;
    }
}

// dst[perm[i]] = src[i] - (sub ? sub[i] : 0)
__global__ void __launch_bounds__(kBlock)
kb_scatter_sub_kernel(int m, const int* __restrict__ perm, const double* __restrict__ src,
                      const double* __restrict__ sub, double* __restrict__ dst) {
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < m; i += gridDim.x * kBlock)
        dst[perm[i]] = src[i] - (sub ? sub[i] : 0.0);
}

// CR right-hand side (:128-141): rhs[k] = free ? 0 : t[k] + a[j] * d.
__global__ void __launch_bounds__(kBlock)
kb_cr_rhs_kernel(int m, const int* __restrict__ basic_var, const double* __restrict__ basic_scale,
                 const unsigned char* __restrict__ free_mask, const double* __restrict__ a,
                 const double* __restrict__ t, double* __restrict__ rhs) {
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < m; k += gridDim.x * kBlock)
        rhs[k] = free_mask[k] ? 0.0 : t[k] + __dmul_rn(a[basic_var[k]], basic_scale[k]);
}

// x[basic_var[k]] = d * t[k]   (:191-193)
__global__ void __launch_bounds__(kBlock)
kb_scatter_basic_kernel(int m, const int* __restrict__ basic_var,
                        const double* __restrict__ basic_scale, const double* __restrict__ t,
                        double* __restrict__ x) {
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < m; k += gridDim.x * kBlock)
        x[basic_var[k]] = __dmul_rn(basic_scale[k], t[k]);
}

// dst[k] = src[perm[k]] * scale[k]
__global__ void __launch_bounds__(kBlock)
kb_gather_scale_kernel(int m, const int* __restrict__ perm, const double* __restrict__ scale,
                       const double* __restrict__ src, double* __restrict__ dst) {
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < m; k += gridDim.x * kBlock)
        dst[k] = __dmul_rn(src[perm[k]], scale[k]);
}

// dst[perm[k]] = src[k] * scale[k]
__global__ void __launch_bounds__(kBlock)
kb_scatter_scale_kernel(int m, const int* __restrict__ perm, const double* __restrict__ scale,
                        const double* __restrict__ src, double* __restrict__ dst) {
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < m; k += gridDim.x * kBlock)
        dst[perm[k]] = __dmul_rn(src[k], scale[k]);
}

// ---- host side ----

static int ensure_reduce(ipxgpu_ctx* c, int grid);
static int launch_normal_apply_w(ipxgpu_ctx* c, const double* Wc, const double* Ws,
                                 const double* x, double* y, int mode, int slot, CrState* st);

static int env_flag(const char* name, int dflt) {
    const char* env = std::getenv(name);
    return env ? std::atoi(env) : dflt;
}

static void free_tri(TriSystem* T) {
    dev_free(T->d.ptr);
    dev_free(T->d.idx);
    dev_free(T->d.val);
    dev_free(T->d.diag);
    dev_free(T->d.order);
    dev_free(T->d.pos);
    dev_free(T->d.level_ptr);
    T->steps.clear();
    T->nlevels = 0;
}

static void destroy_split(ipxgpu_ctx* c) {
    SplitOperator* S = c->split;
    if (!S) return;
    for (int k = 0; k < 4; k++) free_tri(&S->sys[k]);
    dev_free(S->W2);
    dev_free(S->rinv);
    dev_free(S->free_mask);
    dev_free(S->work);
    dev_free(S->xun);
    dev_free(S->yun);
    dev_free(S->basic_var);
    dev_free(S->colperm);
    dev_free(S->basic_scale);
    dev_free(S->tk);
    dev_free(S->wrow);
    delete S;
    c->split = nullptr;
}

static bool split_ready(const ipxgpu_ctx* c) { return c->split && c->split->prepared; }

// Builds one gather-form system from host rows. ascending: rows are solved in
// increasing index order (dependencies have smaller indices).
static int build_tri(ipxgpu_ctx* c, TriSystem* T, int dim, const std::vector<int>& ptr,
                     const std::vector<int>& idx, const std::vector<double>& val,
                     const std::vector<double>* diag, bool ascending, int subtract_seq) {
    free_tri(T);
    T->d.dim = dim;
    T->d.subtract_seq = subtract_seq;
    std::vector<int> level(dim, 0);
    int nlev = dim > 0 ? 1 : 0;
    auto visit = [&](int i) {
        int lv = 0;
        for (int p = ptr[i]; p < ptr[i + 1]; p++) lv = std::max(lv, level[idx[p]] + 1);
        level[i] = lv;
        nlev = std::max(nlev, lv + 1);
    };
    if (ascending) for (int i = 0; i < dim; i++) visit(i);
    else for (int i = dim - 1; i >= 0; i--) visit(i);
    std::vector<int> lptr(nlev + 1, 0);
    for (int i = 0; i < dim; i++) lptr[level[i] + 1]++;
    for (int l = 0; l < nlev; l++) lptr[l + 1] += lptr[l];
    std::vector<int> order(dim);
    {
        std::vector<int> next(lptr.begin(), lptr.end() - 1);
        for (int i = 0; i < dim; i++) order[next[level[i]]++] = i;
    }
    T->nlevels = nlev;
    // Steps: wide levels alone, runs of narrow levels merged.
    int l = 0;
    while (l < nlev) {
        const int width = lptr[l + 1] - lptr[l];
        if (width >= kTriWideRows) {
            T->steps.push_back(TriStep{1, l, l + 1, lptr[l], lptr[l + 1]});
            l++;
            continue;
        }
        int e = l;
        while (e < nlev && lptr[e + 1] - lptr[e] < kTriWideRows) e++;
        T->steps.push_back(TriStep{0, l, e, lptr[l], lptr[e]});
        l = e;
    }
    cudaStream_t s = c->stream;
    IPXGPU_TRY(upload(&T->d.ptr, ptr, s));
    IPXGPU_TRY(upload(&T->d.idx, idx, s));
    IPXGPU_TRY(upload(&T->d.val, val, s));
    if (diag) IPXGPU_TRY(upload(&T->d.diag, *diag, s));
    IPXGPU_TRY(upload(&T->d.order, order, s));
    {
        std::vector<int> pos(dim);
        for (int r = 0; r < dim; r++) pos[order[r]] = r;
        IPXGPU_TRY(upload(&T->d.pos, pos, s));
    }
    IPXGPU_TRY(upload(&T->d.level_ptr, lptr, s));
    IPXGPU_CUDA(cudaStreamSynchronize(s));
    return IPXGPU_OK;
}

static size_t tri_smem_bytes() {
    return (size_t)kTriWarps * kTriStash * sizeof(double) +
           (size_t)kTriWarps * kTriMailDepth * sizeof(ulonglong2);
}

static int launch_tri(ipxgpu_ctx* c, const TriSystem& T, double* x, const CrState* st) {
    static const bool legacy = [] {
        const char* env = std::getenv("IPXGPU_TRI");
        return env && std::string(env) == "levels";
    }();
    if (!legacy) {
        if (T.d.dim == 0) return IPXGPU_OK;
        if (c->tri_grid == 0) {
            const size_t smem = tri_smem_bytes();
            IPXGPU_CUDA(cudaFuncSetAttribute(tri_syncfree_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            IPXGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tri_syncfree_kernel,
                                                                     kTriWarps * 32, smem));
            if (per_sm < 1) return fail(IPXGPU_ERR_STATE, "triangular solve kernel does not fit an SM");
            c->tri_grid = c->num_sms * per_sm;
            IPXGPU_TRY(dev_alloc(&c->tri_err, 1));
            IPXGPU_CUDA(cudaMemsetAsync(c->tri_err, 0, sizeof(unsigned), c->stream));
            IPXGPU_TRY(dev_alloc(&c->tri_ll, (size_t)c->m));
            IPXGPU_CUDA(cudaMemsetAsync(c->tri_ll, 0, sizeof(ulonglong2) * (size_t)c->m, c->stream));
            c->tri_gen = 0;
        }
        if (++c->tri_gen == 0) {  // generation wrapped: start over with clean flags
            IPXGPU_CUDA(cudaMemsetAsync(c->tri_ll, 0, sizeof(ulonglong2) * (size_t)c->m, c->stream));
            c->tri_gen = 1;
        }
        const int grid = std::max(1, std::min(c->tri_grid, (T.d.dim + kTriWarps - 1) / kTriWarps));
        TriDev d = T.d;
        // tuning switches (defaults: on); the summation order of L' rows and of long rows is
        // a context option (ipxgpu_set_option "tri_reference_order")
        static const int sw_prefix = env_flag("IPXGPU_TRI_PREFIX", 1);
        static const int sw_mailbox = env_flag("IPXGPU_TRI_MAILBOX", 1);
        static const int sw_pause = env_flag("IPXGPU_TRI_PAUSE", 32);
        d.prefix = sw_prefix;
        d.mailbox = sw_mailbox;
        d.pause_cap = (unsigned)std::max(0, sw_pause);
        d.trace = c->tri_trace;
        if (d.reverse) d.reverse = c->tri_reference_order ? 1 : 2;
        d.lane_sums = c->tri_reference_order ? 0 : 1;
        double* xp = x;
        ulonglong2* ll = c->tri_ll;
        unsigned gen = c->tri_gen;
        unsigned* err = c->tri_err;
        const CrState* stp = st;
        void* args[] = {&d, &xp, &ll, &gen, &err, &stp};
        IPXGPU_CUDA(cudaLaunchCooperativeKernel((void*)tri_syncfree_kernel, dim3(grid),
                                                dim3(kTriWarps * 32), args, tri_smem_bytes(),
                                                c->stream));
        c->launches++;
        return IPXGPU_OK;
    }
    for (const TriStep& sp : T.steps) {
        if (sp.wide) {
            const int rows = sp.r1 - sp.r0;
            tri_wide_kernel<<<(rows + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(T.d, sp.r0, sp.r1,
                                                                                     x, st);
        } else {
            tri_levels_kernel<<<1, kTriBlock, 0, c->stream>>>(T.d, sp.l0, sp.l1, x, st);
        }
        c->launches++;
    }
    IPXGPU_CUDA(cudaGetLastError());
    return IPXGPU_OK;
}

// After a synchronisation: did a triangular solve give up waiting (tri_row_warp)?
static int check_tri(ipxgpu_ctx* c) {
    if (!c->tri_err) return IPXGPU_OK;
    unsigned e = 0;
    IPXGPU_CUDA(cudaMemcpy(&e, c->tri_err, sizeof e, cudaMemcpyDeviceToHost));
    if (e == 0) return IPXGPU_OK;
    cudaMemset(c->tri_err, 0, sizeof e);
    return fail(IPXGPU_ERR_STATE, "triangular solve: a dependency never resolved");
}

// lhs(m+1) = C*x, lhs[m] = x'lhs (reference src/splitted_normal_matrix.cc:90-117).
static int launch_split_apply(ipxgpu_ctx* c, const double* x, double* lhs, int mode,
                              CrState* st) {
    SplitOperator* S = c->split;
    if (!S || !S->prepared) return fail(IPXGPU_ERR_STATE, "split operator not prepared");
    const int m = (int)c->m;
    const int grid = grid_for(c, m);
    cudaStream_t s = c->stream;
    // work = inverse(B') x : U' then L'
    IPXGPU_CUDA(cudaMemcpyAsync(S->work, x, sizeof(double) * m, cudaMemcpyDeviceToDevice, s));
    IPXGPU_TRY(launch_tri(c, S->sys[2], S->work, st));
    IPXGPU_TRY(launch_tri(c, S->sys[3], S->work, st));
    // lhs = N N' work, through the resident AI with masked squared scales
    gather_perm_kernel<<<grid, kBlock, 0, s>>>(m, S->rinv, S->work, S->xun, st, kSlotBt);
    c->launches++;
    IPXGPU_TRY(launch_normal_apply_w(c, S->W2, S->W2 + c->nloc, S->xun, S->yun, kApplyPlain,
                                     kSlotNone, nullptr));
    scatter_perm_kernel<<<grid, kBlock, 0, s>>>(m, S->rinv, S->yun, lhs, st, kSlotNNt);
    c->launches++;
    // lhs = inverse(B) lhs : L then U
    IPXGPU_TRY(launch_tri(c, S->sys[0], lhs, st));
    IPXGPU_TRY(launch_tri(c, S->sys[1], lhs, st));
    IPXGPU_TRY(ensure_reduce(c, grid));
    split_finish_kernel<<<grid, kBlock, 0, s>>>(m, x, lhs, S->free_mask, c->red, mode, st);
    c->launches++;
    IPXGPU_CUDA(cudaGetLastError());
    return IPXGPU_OK;
}

}  // namespace ipxgpu
