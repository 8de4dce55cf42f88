// Shared-memory tiled gather-reduce sweep.
//
// The generic sweep (spmv_kernels.cuh) gathers the vector through L1/L2. On
// B200 a divergent 8-byte global gather costs ~2 L1 tag cycles per lane, which
// caps a sweep near 20 % of HBM bandwidth (profiles/r01_*). Here the matrix is
// re-tiled in two dimensions at context creation:
//
//   vb  blocks of the GATHER index space (VB entries, staged in shared memory,
//       double-buffered with cp.async)
//   sb  blocks of the SEGMENT space      (SB accumulators in shared memory)
//
// An item = (sb, run of consecutive vb). For every vb of its run the CTA has
// v[vb] in shared memory and walks tile (sb, vb), stored as per-lane streams:
// the segments present in the tile are dealt (longest first) to the lanes of
// the CTA, a lane's stream is the concatenation of its segments' entries, and
// row r of a warp holds the r-th entry of each of its 32 lanes. An entry is a
// 4-byte key (segment16 | last-flag | index15) and an 8-byte value, so a row
// is two fully coalesced loads whose addresses are known in advance: the rows
// of the NEXT tile are prefetched into registers while the current tile is
// processed. A lane gathers from shared memory, sums its segment's entries in
// ascending gather index and, at the segment's last entry, adds the sum to the
// segment's accumulator in shared memory. No shuffles, no atomics, no pointer
// arrays, fixed summation order. When an item covers every vb the epilogue is
// final (t = W .* acc, or y = Ws.*x + acc with the fused dot); otherwise it
// stores a partial vector and tiled_combine_kernel adds the partials in a fixed
// order.
#pragma once

#include <algorithm>
#include <vector>

#include "context.cuh"

namespace ipxgpu {

constexpr int kTiledThreads = 1024;
constexpr int kTiledWarps = kTiledThreads / 32;
constexpr int kTiledVB = 8 * 1024;   // doubles per staged block (64 KB, two buffers)
constexpr int kTiledMaxSB = 8 * 1024;  // accumulators per item (64 KB)
constexpr unsigned kTiledPad = 0xffffu;  // segment id of padding entries
constexpr unsigned kTiledLast = 0x8000u;  // key flag: last entry of its segment in the tile
constexpr int kTiledBatch = 8;            // rows per warp prefetched into registers

struct TiledSweep {
    bool enabled = false;
    int V = 0, S = 0;        // gather-vector length, number of segments
    int VB = 0, SB = 0;      // tile extents
    int NVB = 0, NSB = 0;    // number of blocks
    int K = 0;               // vb blocks per item
    int nparts = 0;          // ceil(NVB / K): partial vectors per segment
    int nitems = 0;          // NSB * nparts
    long long nnz = 0;
    int* row_ptr = nullptr;     // [ntiles * (warps + 1)] first row of each (tile, warp)
    unsigned* keys = nullptr;   // [rows * 32] segment16 << 16 | last << 15 | index15
    double* val = nullptr;      // [rows * 32]
    double* partials = nullptr;       // [nparts * S] when nparts > 1
    size_t smem = 0;
    int debug = 0;  // measurement only: bit 0 skips the staging, bit 1 the entry walk
};

enum TiledMode : int {
    kTiledColScale = 0,   // out[s] = W ? acc*W[s] : acc
    kTiledRowFinal = 1,   // out[s] = (Ws ? x[s]*Ws[s] : 0) + acc, fused dot
    kTiledPartial = 2,    // partials[part*S + s] = acc
};

struct TiledArgs {
    const double* v;       // gather vector
    const double* W;       // kTiledColScale: structural weights or nullptr
    const double* Ws;      // kTiledRowFinal: slack weights or nullptr
    const double* x;       // kTiledRowFinal: rhs for slack term and dot
    double* out;           // t (col mode) or y (row mode, m+1 entries)
    int apply_mode;        // ApplyMode for the fused scalar step
    int slot;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// Starts the asynchronous copy of v[vbase, vbase+vlen) into shared memory
// (16-byte cp.async chunks; plain loads when the base is not 16-byte aligned).
__device__ __forceinline__ void stage_block(double* v_s, const double* v, int vbase, int vlen,
                                            bool aligned) {
    const int tid = threadIdx.x;
    if (aligned) {
        const int pairs = vlen >> 1;
        for (int i = tid; i < pairs; i += kTiledThreads)
            cp_async16(v_s + 2 * i, v + vbase + 2 * i);
        if ((vlen & 1) && tid == 0) v_s[vlen - 1] = v[vbase + vlen - 1];
    } else {
        for (int i = tid; i < vlen; i += kTiledThreads) v_s[i] = v[vbase + i];
    }
    cp_async_commit();
}

__global__ void __launch_bounds__(kTiledThreads, 1)
tiled_sweep_kernel(TiledSweep T, TiledArgs A, int mode, Reduce red, CrState* st) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[32];
    __shared__ int s_flag;
    if (st != nullptr && st->done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // layout: v[2][VB] | acc[SB]
    double* v_buf = reinterpret_cast<double*>(smem_raw);
    double* acc_s = v_buf + 2 * (size_t)T.VB;

    const int sb = blockIdx.x / T.nparts;
    const int part = blockIdx.x - sb * T.nparts;
    const int seg_base = sb * T.SB;
    const int nseg = min(T.SB, T.S - seg_base);
    const int vb0 = part * T.K;
    const int vb1 = min(T.NVB, vb0 + T.K);
    // vb*VB is even, so block starts are 16-byte aligned iff v is.
    const bool aligned = ((reinterpret_cast<unsigned long long>(A.v) & 15ull) == 0);

    for (int s = tid; s < nseg; s += kTiledThreads) acc_s[s] = 0.0;
    {
        const int nvb0 = vb1 - vb0;
        const int first = vb0 + (sb % nvb0);
        stage_block(v_buf, A.v, first * T.VB, min(T.VB, T.V - first * T.VB), aligned);
    }

    const int VBc = T.VB, NVBc = T.NVB;
    // Items of different segment blocks start at different vb blocks so that
    // the SMs do not all pull the same lines of v from L2 at the same time.
    const int nvb = vb1 - vb0;
    const int rot = sb % nvb;
    auto vb_at = [&](int k) { const int r = k + rot; return vb0 + (r >= nvb ? r - nvb : r); };
    // Row ranges of this item's tiles, staged once so that a tile's entry
    // loads do not wait on a dependent pointer load.
    __shared__ int s_rows[64 * (kTiledWarps + 1)];
    const bool rows_in_smem = nvb <= 64;
    if (rows_in_smem) {
        for (int i = tid; i < nvb * (kTiledWarps + 1); i += kTiledThreads)
            s_rows[i] = T.row_ptr[((size_t)sb * NVBc + vb0) * (kTiledWarps + 1) + i];
        __syncthreads();
    }
    const int* row_ptr = T.row_ptr + ((size_t)sb * NVBc) * (kTiledWarps + 1) + warp;

    unsigned kcur[kTiledBatch], knxt[kTiledBatch];
    double acur[kTiledBatch], anxt[kTiledBatch];
    int rb_cur = 0, re_cur = 0, rb_nxt = 0, re_nxt = 0;
    auto fetch = [&](int vb, unsigned* kk, double* aa, int* rb_out, int* re_out) {
        int rb, re;
        if (rows_in_smem) {
            rb = s_rows[(vb - vb0) * (kTiledWarps + 1) + warp];
            re = s_rows[(vb - vb0) * (kTiledWarps + 1) + warp + 1];
        } else {
            rb = __ldg(row_ptr + (size_t)vb * (kTiledWarps + 1));
            re = __ldg(row_ptr + (size_t)vb * (kTiledWarps + 1) + 1);
        }
        *rb_out = rb;
        *re_out = re;
        const unsigned* kp = T.keys + ((size_t)rb << 5) + lane;
        const double* vp = T.val + ((size_t)rb << 5) + lane;
#pragma unroll
        for (int u = 0; u < kTiledBatch; u++) {
            if (rb + u < re) {  // warp-uniform
                kk[u] = __ldcs(kp + (u << 5));
                aa[u] = __ldcs(vp + (u << 5));
            }
        }
    };
    double sum = 0.0;
    auto consume = [&](const double* v_s, unsigned key, double a) {
        sum = sum + __dmul_rn(v_s[key & 0x7fffu], a);
        if (key & kTiledLast) {
            const unsigned seg = key >> 16;
            if (seg != kTiledPad) acc_s[seg] += sum;
            sum = 0.0;
        }
    };
    fetch(vb_at(0), kcur, acur, &rb_cur, &re_cur);
    for (int k = 0; k < nvb; k++) {
        const int cur = k & 1;
        cp_async_wait<0>();
        __syncthreads();  // block k landed; every warp finished the previous tile
        if (k + 1 < nvb) {
            if (!(T.debug & 1)) {
                const int nbase = vb_at(k + 1) * VBc;
                stage_block(v_buf + (size_t)(cur ^ 1) * VBc, A.v, nbase, min(VBc, T.V - nbase),
                            aligned);
            }
            fetch(vb_at(k + 1), knxt, anxt, &rb_nxt, &re_nxt);
        }
        const double* v_s = v_buf + (size_t)cur * VBc;
        if (!(T.debug & 2)) {
#pragma unroll
            for (int u = 0; u < kTiledBatch; u++)
                if (rb_cur + u < re_cur) consume(v_s, kcur[u], acur[u]);
            // tiles with more rows than the register batch: the tail from memory
            for (int r = rb_cur + kTiledBatch; r < re_cur; r++)
                consume(v_s, __ldcs(T.keys + ((size_t)r << 5) + lane),
                        __ldcs(T.val + ((size_t)r << 5) + lane));
        }
#pragma unroll
        for (int u = 0; u < kTiledBatch; u++) {
            kcur[u] = knxt[u];
            acur[u] = anxt[u];
        }
        rb_cur = rb_nxt;
        re_cur = re_nxt;
    }
    __syncthreads();

    double dot = 0.0;
    for (int s = tid; s < nseg; s += kTiledThreads) {
        const int g = seg_base + s;
        const double a = acc_s[s];
        if (mode == kTiledColScale) {
            A.out[g] = A.W ? __dmul_rn(a, A.W[g]) : a;
        } else if (mode == kTiledRowFinal) {
            const double xv = A.x[g];
            const double yv = (A.Ws ? __dmul_rn(xv, A.Ws[g]) : 0.0) + a;
            A.out[g] = yv;
            dot += __dmul_rn(xv, yv);
        } else {
            T.partials[(size_t)part * T.S + g] = a;
        }
    }
    if (mode == kTiledRowFinal) {
        const double mine = block_sum(dot, s_red);
        double ts, ts2, tm;
        if (grid_reduce(red, mine, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) && tid == 0) {
            A.out[T.S] = ts;
            if (st) after_apply(st, A.apply_mode, ts, A.slot);
        }
    }
}

// out = epilogue(sum over parts, in order) for sweeps that ran in partial mode.
__global__ void __launch_bounds__(kBlock)
tiled_combine_kernel(TiledSweep T, TiledArgs A, int mode, Reduce red, CrState* st) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    if (st != nullptr && st->done) return;
    double dot = 0.0;
    for (int g = blockIdx.x * kBlock + threadIdx.x; g < T.S; g += gridDim.x * kBlock) {
        double acc = 0.0;
        for (int p = 0; p < T.nparts; p++) acc += __ldcg(T.partials + (size_t)p * T.S + g);
        if (mode == kTiledColScale) {
            A.out[g] = A.W ? __dmul_rn(acc, A.W[g]) : acc;
        } else {
            const double xv = A.x[g];
            const double yv = (A.Ws ? __dmul_rn(xv, A.Ws[g]) : 0.0) + acc;
            A.out[g] = yv;
            dot += __dmul_rn(xv, yv);
        }
    }
    if (mode == kTiledRowFinal) {
        const double mine = block_sum(dot, s_red);
        double ts, ts2, tm;
        if (grid_reduce(red, mine, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) &&
            threadIdx.x == 0) {
            A.out[T.S] = ts;
            if (st) after_apply(st, A.apply_mode, ts, A.slot);
        }
    }
}

// ---- host side ----

static void free_tiled(TiledSweep* T) {
    dev_free(T->row_ptr);
    dev_free(T->keys);
    dev_free(T->val);
    dev_free(T->partials);
    *T = TiledSweep();
}

// Chooses the tiling for S segments gathering from a vector of length V and
// decides whether it pays: the shared-memory copies of v (read from L2) must
// not exceed `max_ratio` times the matrix stream.
static bool plan_tiled(TiledSweep* T, int V, int S, long long nnz, int num_sms, double max_ratio) {
    if (V <= 0 || S <= 0 || nnz <= 0) return false;
    const int VB = std::min((V + 1) & ~1, kTiledVB);
    const int NVB = (V + VB - 1) / VB;
    // Segment blocks: one per SM if the segments allow.
    int SB = std::min(kTiledMaxSB, std::max(1024, (S + num_sms - 1) / num_sms));
    int NSB = (S + SB - 1) / SB;
    // Every sb row of items stages all of v once.
    auto vbytes = [&](int nsb) { return (double)nsb * V * 8.0; };
    if (vbytes(NSB) > max_ratio * 12.0 * nnz) {
        SB = kTiledMaxSB;  // fewest re-reads
        NSB = (S + SB - 1) / SB;
        if (vbytes(NSB) > max_ratio * 12.0 * nnz) return false;
    }
    // Few segment blocks: spread the vb blocks over the SMs with partial outputs.
    int K = NVB;
    if (NSB < num_sms && NVB > 1) {
        // floor: a single wave of items (a second, nearly empty wave would
        // double the sweep time)
        const int parts = std::min(NVB, std::max(1, num_sms / NSB));
        K = (NVB + parts - 1) / parts;
    }
    const int nparts = (NVB + K - 1) / K;
    T->V = V; T->S = S; T->VB = VB; T->SB = SB; T->NVB = NVB; T->NSB = NSB;
    T->K = K; T->nparts = nparts; T->nitems = NSB * nparts; T->nnz = nnz;
    T->smem = (size_t)2 * VB * 8 + (size_t)SB * 8;
    return true;
}

// Re-tiles a compressed structure (segments ptr[0..S], gather indices idx in
// [0,V), sorted per segment) into T's sliced-ELL layout and uploads it. Returns
// IPXGPU_ERR_UNSUPPORTED (and leaves T disabled) when a segment has more than
// 64 entries inside one vb block (such structures suit the generic sweep).
static int build_tiled(ipxgpu_ctx* c, TiledSweep* T, const int* ptr, const int* idx,
                       const double* val) {
    const int S = T->S, VB = T->VB, SB = T->SB, NVB = T->NVB, NSB = T->NSB;
    const size_t ntiles = (size_t)NSB * NVB;
    struct Run { int seg_local; int first; int len; };
    // pass 1: the runs (segment, first entry, length) of every tile
    std::vector<int> tcount(ntiles + 1, 0);
    for (int s = 0; s < S; s++) {
        const size_t trow = (size_t)(s / SB) * NVB;
        int p = ptr[s];
        const int pe = ptr[s + 1];
        while (p < pe) {
            const int vb = idx[p] / VB;
            int q = p;
            while (q < pe && idx[q] / VB == vb) q++;
            if (q - p > 64) return IPXGPU_ERR_UNSUPPORTED;
            tcount[trow + vb + 1]++;
            p = q;
        }
    }
    for (size_t t = 0; t < ntiles; t++) tcount[t + 1] += tcount[t];
    std::vector<Run> runs((size_t)tcount[ntiles]);
    {
        std::vector<int> next(tcount.begin(), tcount.end() - 1);
        for (int s = 0; s < S; s++) {
            const size_t trow = (size_t)(s / SB) * NVB;
            int p = ptr[s];
            const int pe = ptr[s + 1];
            while (p < pe) {
                const int vb = idx[p] / VB;
                int q = p;
                while (q < pe && idx[q] / VB == vb) q++;
                runs[(size_t)next[trow + vb]++] = Run{s % SB, p, q - p};
                p = q;
            }
        }
    }
    // pass 2: per tile, sort the runs by length (descending, stable) and deal
    // them round-robin to the kTiledThreads lanes; a warp's rows are as many as
    // its longest lane stream.
    const int W = kTiledWarps;
    std::vector<int> row_ptr(ntiles * (size_t)(W + 1), 0);
    std::vector<int> lane_len(kTiledThreads);
    long long rows_total = 0;
    for (size_t t = 0; t < ntiles; t++) {
        Run* rb = runs.data() + tcount[t];
        Run* re = runs.data() + tcount[t + 1];
        std::stable_sort(rb, re, [](const Run& a, const Run& b) { return a.len > b.len; });
        std::fill(lane_len.begin(), lane_len.end(), 0);
        const int nr = (int)(re - rb);
        for (int k = 0; k < nr; k++) lane_len[k % kTiledThreads] += rb[k].len;
        for (int w = 0; w < W; w++) {
            row_ptr[t * (size_t)(W + 1) + w] = (int)rows_total;
            int mx = 0;
            for (int l = 0; l < 32; l++) mx = std::max(mx, lane_len[w * 32 + l]);
            rows_total += mx;
        }
        row_ptr[t * (size_t)(W + 1) + W] = (int)rows_total;
        if (rows_total * 32 >= (long long)INT32_MAX) return IPXGPU_ERR_UNSUPPORTED;
    }
    // pass 3: fill the rows; padding = (kTiledPad, last, index 0, value 0).
    std::vector<unsigned> keys((size_t)rows_total * 32, (kTiledPad << 16) | kTiledLast);
    std::vector<double> vals((size_t)rows_total * 32, 0.0);
    for (size_t t = 0; t < ntiles; t++) {
        const int vb = (int)(t % NVB);
        const Run* rb = runs.data() + tcount[t];
        const int nr = tcount[t + 1] - tcount[t];
        std::fill(lane_len.begin(), lane_len.end(), 0);
        for (int k = 0; k < nr; k++) {
            const int lane_id = k % kTiledThreads;
            const int w = lane_id >> 5, l = lane_id & 31;
            const size_t row0 = (size_t)row_ptr[t * (size_t)(W + 1) + w] + lane_len[lane_id];
            for (int j = 0; j < rb[k].len; j++) {
                const size_t q = (row0 + j) * 32 + l;
                unsigned key = ((unsigned)rb[k].seg_local << 16) |
                               (unsigned)(idx[rb[k].first + j] - vb * VB);
                if (j == rb[k].len - 1) key |= kTiledLast;
                keys[q] = key;
                vals[q] = val[rb[k].first + j];
            }
            lane_len[lane_id] += rb[k].len;
        }
    }
    cudaStream_t s = c->stream;
    IPXGPU_TRY(upload(&T->row_ptr, row_ptr, s));
    IPXGPU_TRY(upload(&T->keys, keys, s));
    IPXGPU_TRY(upload(&T->val, vals, s));
    if (T->nparts > 1) IPXGPU_TRY(dev_alloc(&T->partials, (size_t)T->nparts * S));
    IPXGPU_CUDA(cudaStreamSynchronize(s));
    IPXGPU_CUDA(cudaFuncSetAttribute(tiled_sweep_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(2 * kTiledVB * 8 + kTiledMaxSB * 8)));
    T->enabled = true;
    if (const char* dbg = std::getenv("IPXGPU_TILED_DEBUG")) T->debug = std::atoi(dbg);
    return IPXGPU_OK;
}

static int launch_tiled(ipxgpu_ctx* c, const TiledSweep& T, const TiledArgs& A, int mode,
                        CrState* st) {
    if (T.nparts == 1) {
        tiled_sweep_kernel<<<T.nitems, kTiledThreads, T.smem, c->stream>>>(T, A, mode, c->red, st);
        c->launches++;
    } else {
        tiled_sweep_kernel<<<T.nitems, kTiledThreads, T.smem, c->stream>>>(T, A, kTiledPartial,
                                                                           c->red, st);
        tiled_combine_kernel<<<grid_for(c, T.S), kBlock, 0, c->stream>>>(T, A, mode, c->red, st);
        c->launches += 2;
    }
    IPXGPU_CUDA(cudaGetLastError());
    return IPXGPU_OK;
}

}  // namespace ipxgpu
