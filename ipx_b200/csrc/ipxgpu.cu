// libipxgpu: C ABI over the sm_100a kernels (see include/ipxgpu.h).

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <new>
#include <thread>

#include "context.cuh"
#include "cr_kernels.cuh"
#include "spmv_kernels.cuh"
#include "split.cuh"
#include "band_plan.cuh"
#include "pcr_fused.cuh"
#include "maxvol.cuh"
#include "madd.cuh"

namespace ipxgpu {

thread_local std::string g_last_error;

// ------------------------------------------------------------------ helpers

// Cuts segments [seg_begin, seg_end) of a compressed structure into tiles.
struct HostTiles {
    std::vector<Tile> tiles;
    std::vector<int> long_first;  // prefix over chunks, size num_long+1
};

// `long_chunk`: entries per CTA for segments longer than a tile (the kernel's single-segment
// path takes any length; the spilled dense segments use larger chunks than kTileNnz).
template <class P>
static HostTiles build_tiles(const P* ptr, int seg_begin, int seg_end, long long base,
                             int long_chunk = kTileNnz) {
    HostTiles out;
    out.long_first.push_back(0);
    int s = seg_begin;
    while (s < seg_end) {
        const long long p0 = ptr[s] - base;
        const long long len = ptr[s + 1] - ptr[s];
        if (len > kTileNnz) {
            const int nch = (int)((len + long_chunk - 1) / long_chunk);
            const int long_id = (int)out.long_first.size() - 1;
            for (int c = 0; c < nch; c++) {
                Tile t;
                t.seg0 = s;
                t.nseg = 1;
                t.p0 = (int)(p0 + (long long)c * long_chunk);
                t.p1 = (int)std::min<long long>(p0 + len, p0 + (long long)(c + 1) * long_chunk);
                t.long_id = long_id;
                t.chunk = c;
                out.tiles.push_back(t);
            }
            out.long_first.push_back(out.long_first.back() + nch);
            s++;
            continue;
        }
        int e = s;
        while (e < seg_end && e - s < kTileSeg && (ptr[e + 1] - base) - p0 <= kTileNnz &&
               ptr[e + 1] - ptr[e] <= kTileNnz)
            e++;
        Tile t;
        t.seg0 = s;
        t.nseg = e - s;
        t.p0 = (int)p0;
        t.p1 = (int)(ptr[e] - base);
        t.long_id = -1;
        t.chunk = 0;
        out.tiles.push_back(t);
        s = e;
    }
    return out;
}

static int upload_tiles(TileSet* ts, const HostTiles& h, cudaStream_t s) {
    ts->ntiles = (int)h.tiles.size();
    ts->num_long = (int)h.long_first.size() - 1;
    IPXGPU_TRY(upload(&ts->tiles, h.tiles, s));
    IPXGPU_TRY(upload(&ts->long_first, h.long_first, s));
    IPXGPU_TRY(dev_alloc(&ts->long_partials, (size_t)h.long_first.back()));
    IPXGPU_TRY(dev_alloc(&ts->long_counters, (size_t)ts->num_long));
    IPXGPU_TRY(dev_alloc(&ts->long_dots, (size_t)ts->num_long));
    IPXGPU_CUDA(cudaMemsetAsync(ts->long_dots, 0, sizeof(double) * std::max(1, ts->num_long), s));
    IPXGPU_CUDA(cudaMemsetAsync(ts->long_counters, 0,
                                sizeof(unsigned) * std::max(1, ts->num_long), s));
    return IPXGPU_OK;
}

static void free_tiles(TileSet* ts) {
    dev_free(ts->tiles);
    dev_free(ts->long_first);
    dev_free(ts->long_partials);
    dev_free(ts->long_counters);
    dev_free(ts->long_dots);
}

static void free_matrix(DevMatrix* A) {
    dev_free(A->ptr);
    dev_free(A->idx);
    dev_free(A->val);
}

static int ensure_reduce(ipxgpu_ctx* c, int grid) {
    if (grid <= c->red_cap) return IPXGPU_OK;
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    dev_free(c->red.partials);
    const int cap = std::max(grid, 4096);
    IPXGPU_TRY(dev_alloc(&c->red.partials, (size_t)3 * cap));
    if (!c->red.ticket) {
        IPXGPU_TRY(dev_alloc(&c->red.ticket, 1));
        IPXGPU_CUDA(cudaMemsetAsync(c->red.ticket, 0, sizeof(unsigned), c->stream));
    }
    c->red_cap = cap;
    return IPXGPU_OK;
}

template <class Op>
static int launch_sweep(ipxgpu_ctx* c, const Op& op, const TileSet& ts, const DevMatrix& A,
                        CrState* st) {
    if (ts.ntiles == 0) return IPXGPU_OK;
    seg_sweep_kernel<Op><<<ts.ntiles, kBlock, 0, c->stream>>>(op, ts.tiles, A.ptr, A.idx, A.val,
                                                              ts.info(), c->red, st);
    c->launches++;
    IPXGPU_CUDA(cudaGetLastError());
    return IPXGPU_OK;
}

// ------------------------------------------------------------------ NCCL (dlopen)

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    if (api.handle) return &api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) return nullptr;
    api.handle = h;
    return &api;
}

static int allreduce_sum(ipxgpu_ctx* c, double* buf, size_t count) {
    if (c->nranks == 1) return IPXGPU_OK;
    if (!c->nccl_comm)
        return fail(IPXGPU_ERR_STATE, "nranks > 1 but ipxgpu_comm_init was not called");
    NcclApi* api = nccl_api();
    ncclResult_t r = api->AllReduce(buf, buf, count, ncclDouble, ncclSum,
                                    (ncclComm_t)c->nccl_comm, c->stream);
    if (r != ncclSuccess)
        return fail(IPXGPU_ERR_NCCL, std::string("ncclAllReduce: ") +
                                         (api->GetErrorString ? api->GetErrorString(r) : "?"));
    return IPXGPU_OK;
}

// ------------------------------------------------------------------ operator launches

// Sharded contexts whose ranks exchanged IPC handles sum the partial products with the record
// exchange; IPXGPU_XCHG=nccl (or pull, which exists in the persistent kernel only) keeps
// ncclAllReduce for the launch-per-stage loop.
static bool records_exchange(const ipxgpu_ctx* c) {
    if (c->nranks < 2 || !c->peers_ready || !c->xchg || !c->xchg_abort) return false;
    const char* env = std::getenv("IPXGPU_XCHG");
    if (env && (std::string(env) == "nccl" || std::string(env) == "pull")) return false;
    return true;
}

// lhs(m+1) = AI*W*AI'*x restricted to this shard (allreduced when sharded);
// lhs[m] = x'lhs. `mode`/`slot`/`st` thread the CR scalar step through.
static int launch_normal_apply_w(ipxgpu_ctx* c, const double* Wc, const double* Ws,
                                 const double* x, double* y, int mode, int slot, CrState* st) {
    const int np = (int)c->panels.size();
    const bool sharded = c->nranks > 1;
    if (c->m == 0) return IPXGPU_OK;
    if (np == 0) {
        // no structural columns here: C = diag(W_slack)
        if (sharded) return fail(IPXGPU_ERR_UNSUPPORTED, "a column shard without columns");
        slack_apply_kernel<<<1, kBlock, 0, c->stream>>>((int)c->m, Ws, x, y, mode, slot, st);
        c->launches++;
        IPXGPU_CUDA(cudaGetLastError());
        return IPXGPU_OK;
    }
    const bool band1 = band_usable(c->band1, x), band2 = band_usable(c->band2, c->t);
    if (band1) {
        BandArgs a1{x, Wc, nullptr, nullptr, c->t, kApplyPlain, kSlotNone};
        IPXGPU_TRY(launch_band(c, *c->band1, a1, kBandColScale, st));
        if (c->spill1.nseg > 0) {  // dense columns
            OpColDotScaleSpill op{x, Wc, c->t, c->spill1.map};
            IPXGPU_TRY(launch_sweep(c, op, c->spill1.tiles, c->spill1.A, st));
        }
    }
    for (int k = 0; k < np; k++) {
        const Panel& P = c->panels[k];
        if (!band1) {
            OpColDotScale op1{x, Wc, c->t};
            IPXGPU_TRY(launch_sweep(c, op1, P.col_tiles, c->csc, st));
        }
        if (band2) continue;
        OpRowGather op2;
        op2.t = c->t;
        op2.x = x;
        op2.Ws = (c->rank == 0) ? Ws : nullptr;
        op2.y = y;
        op2.m = (int)c->m;
        op2.first_panel = (k == 0);
        op2.last_panel = (k == np - 1);
        op2.mode = sharded ? (int)kApplyPlain : mode;
        op2.slot = slot;
        IPXGPU_TRY(launch_sweep(c, op2, P.row_tiles, P.csr, sharded ? nullptr : st));
    }
    if (band2) {
        // With spilled (dense) rows the scalar step that follows the apply waits for them.
        const bool spill = c->spill2.nseg > 0;
        const int mode2 = sharded ? (int)kApplyPlain : mode;
        BandArgs a2{c->t, nullptr, (c->rank == 0) ? Ws : nullptr, x, y,
                    spill ? (int)kApplyPlain : mode2, spill ? (int)kSlotNone : slot};
        IPXGPU_TRY(launch_band(c, *c->band2, a2, kBandRowFinal, sharded ? nullptr : st));
        if (spill) {
            OpRowGatherSpill op{c->t, x, y, c->spill2.map, (int)c->m, mode2, slot};
            IPXGPU_TRY(launch_sweep(c, op, c->spill2.tiles, c->spill2.A, sharded ? nullptr : st));
        }
    }
    if (sharded && records_exchange(c)) {
        // Sum over the ranks with the record exchange over NVLink peer memory (pcr_fused.cuh):
        // every rank launches the same sequence of exchanges, numbered by xchg_gen.
        const char* env = std::getenv("IPXGPU_XCHG");
        const std::string how = env ? env : "auto";
        XchgArgs A;
        A.y = y;
        A.x = x;
        A.m = (int)c->m;
        A.nranks = c->nranks;
        A.rank = c->rank;
        A.peers = c->peer_dev;
        A.xmpad = c->xchg_mpad;
        A.xll_off = c->xchg_ll_off;
        A.gen = ++c->xchg_gen;
        A.two_phase = how == "two" || (how != "one" && c->nranks >= 4);
        A.mode = st ? mode : (int)kApplyPlain;
        A.slot = st ? slot : (int)kSlotNone;
        A.abort_word = c->xchg_abort;
        const int grid = grid_for(c, c->m);  // <= 8 CTAs per SM: co-resident
        IPXGPU_TRY(ensure_reduce(c, grid));
        xchg_records_kernel<<<grid, kBlock, 0, c->stream>>>(A, c->red, st);
        c->launches++;
        IPXGPU_CUDA(cudaGetLastError());
    } else if (sharded) {
        IPXGPU_TRY(allreduce_sum(c, y, (size_t)c->m + 1));
        if (st && mode != kApplyPlain) {
            cr_after_apply_kernel<<<1, 1, 0, c->stream>>>(y + c->m, mode, slot, st);
            c->launches++;
        }
    }
    return IPXGPU_OK;
}

static int launch_normal_apply(ipxgpu_ctx* c, const double* x, double* y, int mode, int slot,
                               CrState* st) {
    return launch_normal_apply_w(c, c->Wc, c->Ws, x, y, mode, slot, st);
}

static int launch_diag_build(ipxgpu_ctx* c, const double* Wc, const double* Ws) {
    const int np = (int)c->panels.size();
    for (int k = 0; k < np; k++) {
        const Panel& P = c->panels[k];
        OpRowDiag op{Wc, (c->rank == 0) ? Ws : nullptr, c->diag, k == 0};
        IPXGPU_TRY(launch_sweep(c, op, P.row_tiles, P.csr, nullptr));
    }
    IPXGPU_TRY(allreduce_sum(c, c->diag, (size_t)c->m));
    c->diag_ready = true;
    return IPXGPU_OK;
}

static int launch_operator(ipxgpu_ctx* c, int op, const double* x, double* y, int mode,
                           CrState* st) {
    if (op == 0) return launch_normal_apply(c, x, y, mode, kSlotOp, st);
    return launch_split_apply(c, x, y, mode, st);
}

// ------------------------------------------------------------------ buffers

static int ensure_cr_buffers(ipxgpu_ctx* c, int64_t hist_cap) {
    const size_t m = (size_t)c->m;
    if (!c->v_y) {
        IPXGPU_TRY(dev_alloc(&c->v_y, m));
        IPXGPU_TRY(dev_alloc(&c->v_r, m));
        IPXGPU_TRY(dev_alloc(&c->v_s, m));
        IPXGPU_TRY(dev_alloc(&c->v_p, m));
        IPXGPU_TRY(dev_alloc(&c->v_Cp, m));
        IPXGPU_TRY(dev_alloc(&c->v_Cs, m + 1));
        IPXGPU_TRY(dev_alloc(&c->v_q, m));
        IPXGPU_TRY(dev_alloc(&c->v_rhs, m));
        IPXGPU_TRY(dev_alloc(&c->v_resscale, m));
    }
    if (hist_cap > c->hist_cap) {
        dev_free(c->v_hist);
        IPXGPU_TRY(dev_alloc(&c->v_hist, (size_t)hist_cap));
        c->hist_cap = hist_cap;
    }
    return IPXGPU_OK;
}

static int ensure_nvecs(ipxgpu_ctx* c) {
    for (int k = 0; k < 4; k++)
        if (!c->nvec[k]) IPXGPU_TRY(dev_alloc(&c->nvec[k], (size_t)(c->n + c->m)));
    return IPXGPU_OK;
}

__global__ void cr_start_kernel(CrState* st) { st->t_last = globaltimer(); }

static inline void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}

// ---- persistent CR kernel (pcr_fused.cuh) ----

static bool fused_available(ipxgpu_ctx* c) {
    if ((c->nranks != 1 && !c->peers_ready) || !c->band1 || !c->band2 || c->m <= 0) return false;
    if (c->band1->plan.nparts != 1) return false;  // sweep 1 must write t directly
    if (c->spill1.nseg > 0 || c->spill2.nseg > 0) return false;  // spilled segments: extra launches
    if (const char* env = std::getenv("IPXGPU_FUSED"))
        if (std::atoi(env) == 0) return false;
    return true;
}

static int run_cr_fused(ipxgpu_ctx* c, bool precond, bool zero_start, bool use_resscale,
                        double tol, int64_t maxiter, ipxgpu_cr_result* result,
                        ipxgpu_interrupt_fn interrupt, void* user, int64_t hist_cap) {
    const int m = (int)c->m;
    const char* trace_env = std::getenv("IPXGPU_FUSED_TRACE");
    const bool item_trace = trace_env && std::atoi(trace_env) == 2;
    const bool sharded = c->nranks > 1;
    auto kernel = item_trace ? (sharded ? pcr_fused_kernel<kBandWarps, kBandDepth, 4, 1>
                                        : pcr_fused_kernel<kBandWarps, kBandDepth, 4, 0>)
                             : (sharded ? pcr_fused_kernel<kBandWarps, kBandDepth, 0, 1>
                                        : pcr_fused_kernel<kBandWarps, kBandDepth, 0, 0>);
    const size_t smem = std::max(c->band1->plan.smem, c->band2->plan.smem);
    const int threads = (kBandWarps + 1) * 32;
    IPXGPU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kBandSmemBudget));
    if (c->fused_grid == 0) {
        int per_sm = 0;
        IPXGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
        if (per_sm < 1) return fail(IPXGPU_ERR_STATE, "persistent CR kernel does not fit an SM");
        c->fused_grid = c->num_sms;  // one CTA per SM (shared memory bound)
        IPXGPU_TRY(dev_alloc(&c->fused_bar, 1));
        IPXGPU_TRY(dev_alloc(&c->fused_tickets, (size_t)c->band2->plan.NSB));
        IPXGPU_TRY(dev_alloc(&c->fused_red, (size_t)kFusedStages * 3 * c->fused_grid + 1));
        IPXGPU_TRY(dev_alloc(&c->fused_flags, (size_t)c->band1->plan.nitems + c->fused_grid));
        if (c->fused_grid > kMaxGridSync)
            return fail(IPXGPU_ERR_UNSUPPORTED, "persistent CR kernel: more SMs than kMaxGridSync");
    }
    if (!c->band2->partials)
        IPXGPU_TRY(dev_alloc(&c->band2->partials, (size_t)c->band2->plan.nparts * m));

    CrState h;
    std::memset(&h, 0, sizeof h);
    h.tol = tol;
    h.maxiter = maxiter;
    h.precond = precond ? 1 : 0;
    h.hist = hist_cap > 0 ? c->v_hist : nullptr;
    h.hist_cap = hist_cap;
    h.mirror = c->mirror_dev;
    c->mirror_host->iter = 0;
    c->mirror_host->done = 0;
    c->mirror_host->errflag = 0;
    c->mirror_host->resnorm = 0.0;
    c->mirror_host->abort = 0;
    IPXGPU_CUDA(cudaMemcpyAsync(c->st_dev, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    IPXGPU_CUDA(cudaMemsetAsync(c->fused_red + (size_t)kFusedStages * 3 * c->fused_grid, 0,
                                sizeof(double), c->stream));
    IPXGPU_CUDA(cudaMemsetAsync(c->fused_bar, 0, sizeof(unsigned), c->stream));
    IPXGPU_CUDA(cudaMemsetAsync(c->fused_tickets, 0, sizeof(unsigned) * c->band2->plan.NSB,
                                c->stream));
    IPXGPU_CUDA(cudaMemsetAsync(c->fused_flags, 0,
                                sizeof(unsigned) * ((size_t)c->band1->plan.nitems + c->fused_grid),
                                c->stream));
    if (zero_start) IPXGPU_CUDA(cudaMemsetAsync(c->v_y, 0, sizeof(double) * m, c->stream));

    FusedArgs F;
    F.T1 = *c->band1;
    F.T2 = *c->band2;
    F.v.m = m;
    F.v.y = c->v_y;
    F.v.r = c->v_r;
    F.v.s = c->v_s;
    F.v.p = c->v_p;
    F.v.Cp = c->v_Cp;
    F.v.Cs = c->v_Cs;
    F.v.q = c->v_q;
    F.v.diag = c->diag;
    F.v.resscale = use_resscale ? c->v_resscale : nullptr;
    F.rhs = c->v_rhs;
    F.Wc = c->Wc;
    F.Ws = (c->rank == 0) ? c->Ws : nullptr;  // the slack term is added on one rank
    F.nranks = c->nranks;
    F.rank = c->rank;
    F.peers = c->peer_dev;
    F.xmpad = c->xchg_mpad;
    F.xgen_base = c->xchg_gen;
    {
        const char* env = std::getenv("IPXGPU_XCHG");
        const std::string how = env ? env : "auto";
        F.xll_off = how == "pull" ? 0 : c->xchg_ll_off;
        F.xtwo_phase = how == "two" || (how != "one" && c->nranks >= 4);
        // IPXGPU_XFOLD=0: grid barrier after sweep 2 and the exchange as a stage of its own
        const char* fold = std::getenv("IPXGPU_XFOLD");
        F.xfold = !(fold && std::atoi(fold) == 0);
    }
    F.t = c->t;
    F.zero_start = zero_start ? 1 : 0;
    F.sync = GridSync{c->fused_bar, c->fused_red};
    F.abort_word = c->fused_red + (size_t)kFusedStages * 3 * c->fused_grid;
    F.block_tickets = c->fused_tickets;
    {
        // IPXGPU_FUSED_FLAGS=1: per-piece readiness flags instead of two of the grid barriers
        // of an iteration. Measured on B200 (C2): 82.1 us per apply with the flags against
        // 70.7 us with the barriers - the producer's per-band polls and fences cost more than
        // the 2 x 3 us of barrier they replace - so the barriers stay the default.
        const char* env = std::getenv("IPXGPU_FUSED_FLAGS");
        const bool flags = env && std::atoi(env) != 0;
        F.t_ready = flags ? c->fused_flags : nullptr;
        F.x_ready = flags ? c->fused_flags + c->band1->plan.nitems : nullptr;
    }
    F.st = c->st_dev;
    F.abort_flag = &c->mirror_dev->abort;
    F.trace = nullptr;
    F.trace_cap = 0;
    const bool tracing = trace_env != nullptr;
    const int NWp = 2 * kBandWarps + 4;
    if (item_trace) {
        IPXGPU_TRY(dev_alloc(&F.T1.trace, (size_t)F.T1.plan.nitems * NWp));
        IPXGPU_TRY(dev_alloc(&F.T2.trace, (size_t)F.T2.plan.nitems * NWp));
    }
    if (tracing) {
        F.trace_cap = 4096;
        IPXGPU_TRY(dev_alloc(&F.trace, (size_t)F.trace_cap));
        IPXGPU_CUDA(cudaMemsetAsync(F.trace, 0, 8 * (size_t)F.trace_cap, c->stream));
    }
    void* args[] = {&F};
    IPXGPU_CUDA(cudaLaunchCooperativeKernel((void*)kernel, dim3(c->fused_grid), dim3(threads), args,
                                            smem, c->stream));
    c->launches++;

    int64_t interrupted = 0;
    if (interrupt) {
        // The kernel polls mirror->abort; the host polls the caller's interrupt check.
        for (;;) {
            cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady)
                return fail(IPXGPU_ERR_CUDA, std::string("CR kernel: ") + cudaGetErrorString(q));
            if (!interrupted && (interrupted = interrupt(user)) != 0) c->mirror_host->abort = 1;
            cpu_relax();
        }
    }
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    IPXGPU_CUDA(cudaMemcpy(&h, c->st_dev, sizeof h, cudaMemcpyDeviceToHost));
    if (c->nranks > 1) {
        c->xchg_gen += (unsigned)h.applies;
        double aborted = 0.0;
        IPXGPU_CUDA(cudaMemcpy(&aborted, F.abort_word, sizeof(double), cudaMemcpyDeviceToHost));
        if (aborted == 2.0)
            return fail(IPXGPU_ERR_NCCL, "peer exchange timed out: a rank did not reach the CR solve");
    }
    if (tracing) {
        // CTA 0, 10 stamps per iteration after the initialisation:
        // dir | dir sync | upd | upd sync | s1 | s1 sync | s2 | s2 sync | comb | comb sync
        std::vector<unsigned long long> tr((size_t)F.trace_cap);
        cudaMemcpy(tr.data(), F.trace, 8 * tr.size(), cudaMemcpyDeviceToHost);
        int n = 0;
        while (n < F.trace_cap && tr[n]) n++;
        fprintf(stderr, "[fused trace] %d stamps; deltas (us):", n);
        for (int k = 1; k < n && k < 80; k++) fprintf(stderr, " %.2f", (tr[k] - tr[k - 1]) / 1e3);
        fprintf(stderr, "\n");
        cudaFree(F.trace);
    }
    if (item_trace) {
        // per-item stamps of the LAST apply: start | prologue end | warps done | end | warp ends
        for (int sw = 1; sw <= 2; sw++) {
            const BandDev& T = sw == 1 ? F.T1 : F.T2;
            const int ni = T.plan.nitems;
            std::vector<unsigned long long> tr((size_t)ni * NWp);
            cudaMemcpy(tr.data(), T.trace, 8 * tr.size(), cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ull;
            for (int i = 0; i < ni; i++) t0 = std::min(t0, tr[(size_t)i * NWp]);
            const char* names[4] = {"start (rel)", "prologue", "main", "epilogue"};
            for (int q = 0; q < 4; q++) {
                double mn = 1e30, mx = 0, sum = 0;
                for (int i = 0; i < ni; i++) {
                    const unsigned long long* r = tr.data() + (size_t)i * NWp;
                    const double v = q == 0 ? (double)(r[0] - t0) : (double)(r[q] - r[q - 1]);
                    mn = std::min(mn, v);
                    mx = std::max(mx, v);
                    sum += v;
                }
                fprintf(stderr, "[fused item trace] sweep %d %-12s min %6.2f mean %6.2f max %6.2f us\n",
                        sw, names[q], mn / 1e3, sum / ni / 1e3, mx / 1e3);
            }
            cudaFree(T.trace);
        }
    }
    if (result) {
        result->errflag = interrupted ? interrupted : h.errflag;
        result->iter = h.iter;
        result->time_op = 1e-9 * (double)h.t_op;
        result->time_pre = 1e-9 * (double)h.t_pre;
        result->time_B = 0.0;
        result->time_Bt = 0.0;
        result->time_NNt = 0.0;
        result->resnorm = h.resnorm;
    }
    return IPXGPU_OK;
}

// Device-resident CR driver. Vectors v_rhs, v_y (initial iterate, only if
// !zero_start), v_resscale (if use_resscale) must be on the device already.
// z = inv(S) Ad' inv(E) v for one or two vectors (smw.cuh), on the solve's stream.
static int launch_smw_pass(ipxgpu_ctx* c, const double* v0, const double* v1, int which,
                           const CrState* st) {
    const SmwDev& S = c->smw;
    smw_gather_kernel<<<S.nchunks, kBlock, 0, c->stream>>>(S, c->diag, v0, v1, which, st);
    smw_solve_kernel<<<1, kSmwSolveThreads, 0, c->stream>>>(S, which, st);
    c->launches += 2;
    IPXGPU_CUDA(cudaGetLastError());
    return IPXGPU_OK;
}

static int smw_grid(const ipxgpu_ctx* c) {
    return c->smw.warp_rows ? grid_for(c, (long long)c->m * 32) : grid_for(c, c->m);
}

static int run_cr(ipxgpu_ctx* c, int op, bool precond, bool zero_start, bool use_resscale,
                  double tol, int64_t maxiter, ipxgpu_cr_result* result,
                  ipxgpu_interrupt_fn interrupt, void* user, int64_t hist_cap) {
    const int m = (int)c->m;
    if (maxiter < 0) maxiter = c->m + 100;
    const bool smw = precond && c->smw_active;  // preconditioner with a dense-column part
    if (op == 0 && fused_available(c) && !smw)
        return run_cr_fused(c, precond, zero_start, use_resscale, tol, maxiter, result, interrupt,
                            user, hist_cap);
    const int grid = grid_for(c, m);
    IPXGPU_TRY(ensure_reduce(c, grid));

    CrState h;
    std::memset(&h, 0, sizeof h);
    h.tol = tol;
    h.maxiter = maxiter;
    h.precond = precond ? 1 : 0;
    h.hist = hist_cap > 0 ? c->v_hist : nullptr;
    h.hist_cap = hist_cap;
    h.mirror = c->mirror_dev;
    c->mirror_host->iter = 0;
    c->mirror_host->done = 0;
    c->mirror_host->errflag = 0;
    c->mirror_host->resnorm = 0.0;
    IPXGPU_CUDA(cudaMemcpyAsync(c->st_dev, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    cr_start_kernel<<<1, 1, 0, c->stream>>>(c->st_dev);
    c->launches++;

    CrVectors v;
    v.m = m;
    v.y = c->v_y;
    v.r = c->v_r;
    v.s = c->v_s;
    v.p = c->v_p;
    v.Cp = c->v_Cp;
    v.Cs = c->v_Cs;
    v.q = c->v_q;
    v.diag = c->diag;
    v.resscale = use_resscale ? c->v_resscale : nullptr;
    CrState* st = c->st_dev;

    if (m == 0) {
        if (result) std::memset(result, 0, sizeof *result);
        return IPXGPU_OK;
    }

    // Initialisation (reference src/conjugate_residuals.cc:33-40, :118-127).
    if (zero_start) {
        IPXGPU_CUDA(cudaMemsetAsync(c->v_y, 0, sizeof(double) * m, c->stream));
        cr_init_kernel<<<grid, kBlock, 0, c->stream>>>(v, c->v_rhs, nullptr, c->red, st);
    } else {
        IPXGPU_TRY(launch_operator(c, op, c->v_y, c->v_Cs, kApplyPlain, nullptr));
        cr_init_kernel<<<grid, kBlock, 0, c->stream>>>(v, c->v_rhs, c->v_Cs, c->red, st);
    }
    c->launches++;
    const int sgrid = smw ? smw_grid(c) : 0;
    if (smw) {
        // s = inv(P) r with the dense-column part (cr_init_kernel's s = r ./ E is replaced)
        IPXGPU_TRY(ensure_reduce(c, std::max(grid, sgrid)));
        IPXGPU_TRY(launch_smw_pass(c, c->v_r, nullptr, kSmwFirst, nullptr));
        cr_init_smw_kernel<<<sgrid, kBlock, 0, c->stream>>>(v, c->smw, c->red, st);
        c->launches++;
    }
    // Direction stage: p, Cp, q = P Cp, the dots and the tests; with a dense-column part the
    // preconditioner apply needs the complete Cp first.
    auto direction = [&]() -> int {
        cr_direction_kernel<<<grid, kBlock, 0, c->stream>>>(v, c->red, st, smw ? 1 : 0);
        c->launches++;
        if (smw) {
            IPXGPU_TRY(launch_smw_pass(c, c->v_Cp, c->v_r, kSmwRecompute, st));
            cr_direction_smw_kernel<<<sgrid, kBlock, 0, c->stream>>>(v, c->smw, c->red, st);
            c->launches++;
        }
        return IPXGPU_OK;
    };
    IPXGPU_TRY(launch_operator(c, op, precond ? c->v_s : c->v_r, c->v_Cs, kApplyCrInit, st));
    IPXGPU_TRY(direction());
    IPXGPU_CUDA(cudaGetLastError());

    auto enqueue_pass = [&]() -> int {
        cr_update_kernel<<<grid, kBlock, 0, c->stream>>>(v, c->red, st);
        c->launches++;
        IPXGPU_TRY(launch_operator(c, op, precond ? c->v_s : c->v_r, c->v_Cs, kApplyCrIter, st));
        return direction();
    };

    const int kBatch = 8;
    long long enqueued = 0;
    int64_t interrupted = 0;
    volatile HostMirror* mir = c->mirror_host;
    if (c->nranks > 1) {
        // All ranks must issue the same collectives: advance in lock step.
        for (;;) {
            IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
            if (mir->done) break;
            if (interrupt && (interrupted = interrupt(user)) != 0) break;
            for (int b = 0; b < kBatch; b++) IPXGPU_TRY(enqueue_pass());
        }
    } else {
        unsigned long long polls = 0;
        for (;;) {
            if (mir->done) break;
            if (enqueued - mir->iter <= kBatch) {
                if (interrupt && (interrupted = interrupt(user)) != 0) break;
                for (int b = 0; b < kBatch; b++) IPXGPU_TRY(enqueue_pass());
                enqueued += kBatch;
                IPXGPU_CUDA(cudaGetLastError());
                continue;
            }
            if ((++polls & 0xfff) == 0) {
                cudaError_t q = cudaStreamQuery(c->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady)
                    return fail(IPXGPU_ERR_CUDA, std::string("CR loop: ") + cudaGetErrorString(q));
                if (q == cudaSuccess && !mir->done && enqueued - mir->iter > kBatch)
                    return fail(IPXGPU_ERR_STATE, "CR loop stalled: stream idle but not done");
            }
            cpu_relax();
        }
    }
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    IPXGPU_CUDA(cudaMemcpy(&h, c->st_dev, sizeof h, cudaMemcpyDeviceToHost));
    if (c->nranks > 1 && c->xchg_abort) {
        double aborted = 0.0;
        IPXGPU_CUDA(cudaMemcpy(&aborted, c->xchg_abort, sizeof aborted, cudaMemcpyDeviceToHost));
        if (aborted != 0.0) {
            cudaMemset(c->xchg_abort, 0, sizeof(double));
            return fail(IPXGPU_ERR_STATE, "peer exchange timed out: a rank did not arrive");
        }
    }
    if (result) {
        result->errflag = interrupted ? interrupted : h.errflag;
        result->iter = h.iter;
        result->time_op = 1e-9 * (double)h.t_op;
        result->time_pre = 1e-9 * (double)h.t_pre;
        result->time_B = 1e-9 * (double)h.t_B;
        result->time_Bt = 1e-9 * (double)h.t_Bt;
        result->time_NNt = 1e-9 * (double)h.t_NNt;
        if (op == 1) result->time_op = result->time_B + result->time_Bt + result->time_NNt;
        result->resnorm = h.resnorm;
    }
    return IPXGPU_OK;
}

static int check_ctx(const ipxgpu_ctx* c) {
    if (!c) return fail(IPXGPU_ERR_ARGUMENT, "null context");
    if (c->group)
        return fail(IPXGPU_ERR_UNSUPPORTED,
                    "a multi-GPU group drives the KKTSolverDiag entry points only "
                    "(ipxgpu_kktdiag_factorize / ipxgpu_kktdiag_solve)");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(IPXGPU_ERR_CUDA, cudaGetErrorString(e));
    return IPXGPU_OK;
}

}  // namespace ipxgpu

using namespace ipxgpu;

// The contexts of a one-process, several-GPU group (see ipxgpu_create_group).
struct ipxgpu_group {
    std::vector<ipxgpu_ctx*> sub;
};

// =================================================================== C ABI

extern "C" {

void ipxgpu_default_options(ipxgpu_options* opt) {
    if (!opt) return;
    opt->device = -1;
    opt->rank = 0;
    opt->nranks = 1;
    opt->col_begin = -1;
    opt->col_end = -1;
    opt->panel_cols = 0;
    opt->stream = nullptr;
}

const char* ipxgpu_last_error(void) { return g_last_error.c_str(); }

// CUDA runtime and device context creation take 0.7-5 s in a fresh process. It runs on a helper
// thread so that host-side work (model loading, layout builds) overlaps it; whoever needs the
// device first joins the thread.
namespace {
struct Warmup {
    std::mutex mu;
    std::thread th;
    bool started = false, joined = false;
    int device = -1, ndev = 0;
    cudaError_t err = cudaSuccess;
    double ms = 0.0;
    void start(int dev_hint) {
        std::lock_guard<std::mutex> lock(mu);
        if (started) return;
        started = true;
        th = std::thread([this, dev_hint] {
            const auto t0 = std::chrono::steady_clock::now();
            err = cudaGetDeviceCount(&ndev);
            if (err == cudaSuccess && ndev > 0) {
                // Without a device hint only the driver is initialised here: the caller's
                // current device is a property of ITS thread, resolved in ipxgpu_create.
                int dev = dev_hint;
                if (dev < 0) {
                    const char* env = std::getenv("IPXGPU_DEVICE");
                    if (env) dev = std::atoi(env);
                }
                device = dev;
                if (dev >= 0 && dev < ndev) {
                    err = cudaSetDevice(dev);
                    if (err == cudaSuccess) err = cudaFree(nullptr);  // creates the context
                }
            }
            ms = 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        });
    }
    void wait() {
        std::lock_guard<std::mutex> lock(mu);
        if (started && !joined) {
            if (th.joinable()) th.join();
            joined = true;
        }
    }
};
// Never destroyed: a process that exits while the helper thread is still inside the CUDA
// initialisation must neither wait for it nor run a std::thread destructor on it.
Warmup& warmup() {
    static Warmup* w = new Warmup;
    return *w;
}
}  // namespace

int ipxgpu_warmup(int32_t device) {
    warmup().start(device);
    return IPXGPU_OK;
}

int ipxgpu_device_count(int* count) {
    if (!count) return fail(IPXGPU_ERR_ARGUMENT, "null count");
    *count = 0;
    IPXGPU_CUDA(cudaGetDeviceCount(count));
    return IPXGPU_OK;
}

int ipxgpu_partition_columns(int64_t n, const int64_t* AIp, int32_t nranks, int64_t* bounds) {
    if (n < 0 || !AIp || nranks < 1 || !bounds)
        return fail(IPXGPU_ERR_ARGUMENT, "invalid partition arguments");
    const int64_t nnzA = AIp[n] - AIp[0];
    bounds[0] = 0;
    for (int r = 1; r < nranks; r++) {
        const int64_t target = AIp[0] + (int64_t)((__int128)nnzA * r / nranks);
        int64_t cut = std::lower_bound(AIp, AIp + n + 1, target) - AIp;
        if (cut > n) cut = n;
        bounds[r] = std::max(cut, bounds[r - 1]);
    }
    bounds[nranks] = n;
    return IPXGPU_OK;
}

static void free_smw(ipxgpu_ctx* c);

static void destroy_group(ipxgpu_ctx* c);
static int group_kktdiag_factorize(ipxgpu_ctx* c, const double* xl, const double* xu,
                                   const double* zl, const double* zu, double mu, double* W_out,
                                   double* resscale_out);
static int group_kktdiag_solve(ipxgpu_ctx* c, const double* a, const double* b, double tol,
                               int64_t maxiter, double* x, double* y, ipxgpu_cr_result* result,
                               ipxgpu_interrupt_fn interrupt, void* user);

void ipxgpu_destroy(ipxgpu_ctx* c) {
    if (!c) return;
    if (c->group) {
        destroy_group(c);
        return;
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    destroy_split(c);
    dev_free(c->mv_colscale); dev_free(c->mv_colweights); dev_free(c->mv_vec);
    dev_free(c->mv_partials); dev_free(c->mv_out); dev_free(c->mv_ticket);
    for (int r = 0; r < c->nranks && r < 16; r++)
        if (r != c->rank && c->peer_base[r] && !c->peers_direct)
            cudaIpcCloseMemHandle(c->peer_base[r]);
    if (c->xchg) cudaFree(c->xchg);
    dev_free(c->peer_dev);
    dev_free(c->fused_bar);
    dev_free(c->tri_ll);
    dev_free(c->tri_trace);
    dev_free(c->xchg_abort);
    dev_free(c->tri_err);
    dev_free(c->fused_tickets);
    dev_free(c->fused_flags);
    free_smw(c);
    dev_free(c->W_mask);
    dev_free(c->fused_red);
    if (c->band1) { free_band(c->band1); delete c->band1; }
    if (c->band2) { free_band(c->band2); delete c->band2; }
    for (Spill* sp : {&c->spill1, &c->spill2}) {
        free_matrix(&sp->A);
        free_tiles(&sp->tiles);
        dev_free(sp->map);
    }
    free_matrix(&c->csc);
    for (Panel& P : c->panels) {
        free_tiles(&P.col_tiles);
        free_tiles(&P.row_tiles);
        free_matrix(&P.csr);
    }
    dev_free(c->W_own);
    dev_free(c->W_full);
    dev_free(c->resscale_kkt);
    dev_free(c->t);
    dev_free(c->xin);
    dev_free(c->ybuf);
    dev_free(c->diag);
    dev_free(c->v_y); dev_free(c->v_r); dev_free(c->v_s); dev_free(c->v_p); dev_free(c->v_Cp);
    dev_free(c->v_Cs); dev_free(c->v_q); dev_free(c->v_rhs); dev_free(c->v_resscale);
    dev_free(c->v_hist);
    for (int k = 0; k < 4; k++) dev_free(c->nvec[k]);
    dev_free(c->red.partials);
    dev_free(c->red.ticket);
    dev_free(c->st_dev);
    dev_free(c->scalars);
    dev_free(c->flush_buf);
    if (c->mirror_host) cudaFreeHost(c->mirror_host);
    if (c->nccl_comm) {
        NcclApi* api = nccl_api();
        if (api) api->CommDestroy((ncclComm_t)c->nccl_comm);
    }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int ipxgpu_create(ipxgpu_ctx** out, int64_t m, int64_t n, const int64_t* AIp, const int64_t* AIi,
                  const double* AIx, const ipxgpu_options* opt_in) {
    if (!out) return fail(IPXGPU_ERR_ARGUMENT, "null ctx out");
    *out = nullptr;
    if (m < 0 || n < 0 || !AIp || (AIp[n] > 0 && (!AIi || !AIx)))
        return fail(IPXGPU_ERR_ARGUMENT, "invalid matrix arguments");
    ipxgpu_options opt;
    ipxgpu_default_options(&opt);
    if (opt_in) opt = *opt_in;
    if (opt.nranks < 1 || opt.rank < 0 || opt.rank >= opt.nranks)
        return fail(IPXGPU_ERR_ARGUMENT, "invalid rank/nranks");

    // ---- host-only part: runs while the helper thread creates the CUDA context ----
    const auto t_enter = std::chrono::steady_clock::now();
    warmup().start(opt.device);  // no-op when ipxgpu_warmup was called earlier in the process
    const bool timing = std::getenv("IPXGPU_TIMING") != nullptr;
    auto t_last = t_enter;
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ipxgpu_create] %-34s %8.1f ms\n", what,
                1e3 * std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };

    // Column shard: explicit, or balanced by nonzeros.
    int64_t cb = opt.col_begin, ce = opt.col_end;
    if (cb < 0 || ce < 0) {
        std::vector<int64_t> bounds((size_t)opt.nranks + 1);
        IPXGPU_TRY(ipxgpu_partition_columns(n, AIp, opt.nranks, bounds.data()));
        cb = bounds[opt.rank];
        ce = bounds[opt.rank + 1];
    }
    if (cb < 0 || ce < cb || ce > n) return fail(IPXGPU_ERR_ARGUMENT, "invalid column range");
    const int64_t nloc = ce - cb;
    const int64_t base = AIp[cb];
    const int64_t nnz = AIp[ce] - base;
    if (nnz >= (int64_t)INT32_MAX || nloc >= (int64_t)INT32_MAX || m >= (int64_t)INT32_MAX)
        return fail(IPXGPU_ERR_UNSUPPORTED, "shard exceeds int32 index range; use more shards");

    // The band layouts are planned for the SM count of a B200 while the device is not known
    // yet, and re-planned in the (not expected) case that the device reports another count.
    const char* env_sms = std::getenv("IPXGPU_NUM_SMS");
    const int spec_sms = env_sms ? std::max(1, std::atoi(env_sms)) : 148;
    const char* env_sweep = std::getenv("IPXGPU_SWEEP");
    const std::string how = env_sweep ? env_sweep : "auto";
    const bool force = how == "band" || how == "tiled";
    const double ratio = force ? 1e30 : 1.5;
    const double max_pad = force ? 1.0 : 0.25;
    const bool want_band = how != "generic" && nnz > 0 && m > 0;

    int64_t pc = opt.panel_cols > 0 ? opt.panel_cols : (int64_t)2 << 20;
    const char* env_pc = std::getenv("IPXGPU_PANEL_COLS");
    if (opt.panel_cols <= 0 && env_pc) pc = std::max<int64_t>(1, std::atoll(env_pc));
    const int npanels = (int)std::max<int64_t>(1, (nloc + pc - 1) / pc);

    std::vector<int> cp, ci, rp, rj;  // shard CSC (int32) and, when needed, the full-shard CSR
    std::vector<double> rx;
    struct BandJob {  // declared after the vectors its thread reads: joined before they die
        BandPlan plan;
        BandHost H;
        std::vector<int> spilled;  // segments left to the generic kernel
        bool planned = false, built = false, oom = false;
        std::thread th;
    } job1, job2;
    struct Joiner {
        BandJob *a, *b;
        ~Joiner() {
            if (a->th.joinable()) a->th.join();
            if (b->th.joinable()) b->th.join();
        }
    } joiner{&job1, &job2};

    bool have_csr = false;
    try {
        cp.resize((size_t)nloc + 1);
        ci.resize((size_t)nnz);
        for (int64_t j = 0; j <= nloc; j++) cp[j] = (int)(AIp[cb + j] - base);
        for (int64_t p = 0; p < nnz; p++) {
            const int64_t i = AIi[base + p];
            if (i < 0 || i >= m) return fail(IPXGPU_ERR_ARGUMENT, "row index out of range");
            ci[p] = (int)i;
        }
        lap("host: CSC int32");
        auto run_job = [&](BandJob* job, const int* ptr, const int* idx, const double* val) {
            job->th = std::thread([=] {
                try {
                    job->built = band_build(job->plan, ptr, idx, val, &job->H, kBandMaxRun,
                                            &job->spilled);
                } catch (const std::bad_alloc&) {
                    job->oom = true;
                }
            });
        };
        if (want_band) {
            job1.planned = plan_band(&job1.plan, (int)m, (int)nloc, nnz, spec_sms, ratio);
            job2.planned = plan_band(&job2.plan, (int)nloc, (int)m, nnz, spec_sms, ratio);
        }
        if (job1.planned) run_job(&job1, cp.data(), ci.data(), AIx + base);
        if (job2.planned || npanels == 1) {
            // full-shard CSR (rows ascending, columns ascending within a row)
            rp.assign((size_t)m + 1, 0);
            rj.resize((size_t)nnz);
            rx.resize((size_t)nnz);
            for (int64_t p = 0; p < nnz; p++) rp[ci[p] + 1]++;
            for (int64_t i = 0; i < m; i++) rp[i + 1] += rp[i];
            std::vector<int> next(rp.begin(), rp.end() - 1);
            for (int j = 0; j < (int)nloc; j++)
                for (int p = cp[j]; p < cp[j + 1]; p++) {
                    const int put = next[ci[p]]++;
                    rj[put] = j;
                    rx[put] = AIx[base + p];
                }
            have_csr = true;
            lap("host: CSR of the shard");
            if (job2.planned) run_job(&job2, rp.data(), rj.data(), rx.data());
        }
    } catch (const std::bad_alloc&) {
        return fail(IPXGPU_ERR_OUT_OF_MEMORY, "host allocation failed while building layouts");
    }

    // ---- device part ----
    Warmup& wu = warmup();
    wu.wait();
    lap("wait for CUDA runtime / context");
    if (timing)
        fprintf(stderr, "[ipxgpu_create] %-34s %8.1f ms (helper thread)\n",
                "CUDA runtime / context creation", wu.ms);
    if (wu.err != cudaSuccess || wu.ndev == 0)
        return fail(IPXGPU_ERR_CUDA, std::string("no CUDA device (libipxgpu has no CPU path): ") +
                                         cudaGetErrorString(wu.err));
    const int ndev = wu.ndev;
    int device = opt.device;
    if (device < 0) {
        const char* env = std::getenv("IPXGPU_DEVICE");
        if (env) device = std::atoi(env);
        else IPXGPU_CUDA(cudaGetDevice(&device));
    }
    if (device < 0 || device >= ndev) return fail(IPXGPU_ERR_ARGUMENT, "device ordinal out of range");
    IPXGPU_CUDA(cudaSetDevice(device));
    IPXGPU_CUDA(cudaFree(nullptr));  // this device's context, if the helper warmed another one

    ipxgpu_ctx* c = new (std::nothrow) ipxgpu_ctx;
    if (!c) return fail(IPXGPU_ERR_OUT_OF_MEMORY, "host allocation failed");
    struct Guard {
        ipxgpu_ctx* c;
        ~Guard() { if (c) ipxgpu_destroy(c); }
    } guard{c};

    c->device = device;
    c->m = m;
    c->n = n;
    c->rank = opt.rank;
    c->nranks = opt.nranks;
    cudaDeviceProp prop;
    IPXGPU_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    if (const char* env = std::getenv("IPXGPU_TRI_ORDER"))
        c->tri_reference_order = std::string(env) == "reference";
    if (opt.stream) {
        c->stream = (cudaStream_t)opt.stream;
    } else {
        IPXGPU_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->col_begin = cb;
    c->col_end = ce;
    c->nloc = (int)nloc;
    c->csc.nseg = (int)nloc;
    c->csc.nnz = nnz;

    cudaStream_t s = c->stream;
    try {
        IPXGPU_TRY(upload(&c->csc.ptr, cp, s));
        IPXGPU_TRY(upload(&c->csc.idx, ci, s));
        IPXGPU_TRY(dev_alloc(&c->csc.val, (size_t)nnz));
        if (nnz > 0)
            IPXGPU_CUDA(cudaMemcpyAsync(c->csc.val, AIx + base, sizeof(double) * nnz,
                                        cudaMemcpyHostToDevice, s));
        lap("upload CSC");
        // Panels: t-slice of at most panel_cols columns (default 2M = 16 MB).
        // A shard without structural columns still gets one (empty) panel: its
        // row sweep carries the slack term and the fused dot product.
        c->panels.resize(npanels);
        int max_grid = 1;
        for (int k = 0; k < (int)c->panels.size(); k++) {
            Panel& P = c->panels[k];
            P.c0 = (int)(nloc * k / npanels);
            P.c1 = (int)(nloc * (k + 1) / npanels);
            HostTiles ct = build_tiles(cp.data(), P.c0, P.c1, 0);
            IPXGPU_TRY(upload_tiles(&P.col_tiles, ct, s));
            // CSR of the panel (rows global, columns local to the shard,
            // ascending within a row).
            const int64_t q0 = cp[P.c0], q1 = cp[P.c1];
            std::vector<int> prp, prj;
            std::vector<double> prx;
            const bool whole = npanels == 1 && have_csr;
            if (!whole) {
                prp.assign((size_t)m + 1, 0);
                for (int64_t p = q0; p < q1; p++) prp[ci[p] + 1]++;
                for (int64_t i = 0; i < m; i++) prp[i + 1] += prp[i];
                prj.resize((size_t)(q1 - q0));
                prx.resize((size_t)(q1 - q0));
                std::vector<int> next(prp.begin(), prp.end() - 1);
                for (int j = P.c0; j < P.c1; j++)
                    for (int p = cp[j]; p < cp[j + 1]; p++) {
                        const int put = next[ci[p]]++;
                        prj[put] = j;
                        prx[put] = AIx[base + p];
                    }
            }
            const std::vector<int>& urp = whole ? rp : prp;
            const std::vector<int>& urj = whole ? rj : prj;
            const std::vector<double>& urx = whole ? rx : prx;
            P.csr.nseg = (int)m;
            P.csr.nnz = q1 - q0;
            IPXGPU_TRY(upload(&P.csr.ptr, urp, s));
            IPXGPU_TRY(upload(&P.csr.idx, urj, s));
            IPXGPU_TRY(upload(&P.csr.val, urx, s));
            HostTiles rt = build_tiles(urp.data(), 0, (int)m, 0);
            IPXGPU_TRY(upload_tiles(&P.row_tiles, rt, s));
            max_grid = std::max(max_grid, std::max(P.col_tiles.ntiles, P.row_tiles.ntiles));
            IPXGPU_CUDA(cudaStreamSynchronize(s));  // host vectors die here
        }
        lap("panels: CSR, tiles, upload");
        // Banded shared-memory sweeps (band_sweep.cuh) carry the normal-matrix
        // apply wherever the structure suits them. IPXGPU_SWEEP=generic turns
        // them off, IPXGPU_SWEEP=band forces them whenever a plan fits.
        if (job1.th.joinable()) job1.th.join();
        if (job2.th.joinable()) job2.th.join();
        lap("wait for the band layouts");
        if (job1.oom || job2.oom)
            return fail(IPXGPU_ERR_OUT_OF_MEMORY, "host allocation failed while building layouts");
        if (want_band && c->num_sms != spec_sms) {  // not a B200: plan again for this device
            for (int k = 0; k < 2; k++) {
                BandJob& job = k == 0 ? job1 : job2;
                job.H = BandHost();
                job.built = false;
                job.planned = k == 0 ? plan_band(&job.plan, (int)m, (int)nloc, nnz, c->num_sms, ratio)
                                     : plan_band(&job.plan, (int)nloc, (int)m, nnz, c->num_sms, ratio);
                if (!job.planned || (k == 1 && !have_csr)) continue;
                job.built = k == 0 ? band_build(job.plan, cp.data(), ci.data(), AIx + base, &job.H,
                                                kBandMaxRun, &job.spilled)
                                   : band_build(job.plan, rp.data(), rj.data(), rx.data(), &job.H,
                                                kBandMaxRun, &job.spilled);
            }
            lap("band layouts re-planned");
        }
        for (int k = 0; k < 2; k++) {
            BandJob& job = k == 0 ? job1 : job2;
            if (!job.planned || !job.built) continue;  // structure suits the generic sweep
            BandDev* T = new BandDev();
            T->plan = job.plan;
            const int rc = upload_band(c, T, job.H, max_pad);
            job.H = BandHost();
            if (rc != IPXGPU_OK) {
                free_band(T);
                delete T;
                if (rc == IPXGPU_ERR_UNSUPPORTED) continue;
                return rc;
            }
            (k == 0 ? c->band1 : c->band2) = T;
            max_grid = std::max(max_grid, job.plan.nitems);
            if (!job.spilled.empty()) {
                // compact copy of the spilled segments for the generic kernel
                const int* ptr = k == 0 ? cp.data() : rp.data();
                const int* idx = k == 0 ? ci.data() : rj.data();
                const double* val = k == 0 ? AIx + base : rx.data();
                Spill& sp = k == 0 ? c->spill1 : c->spill2;
                const int ns = (int)job.spilled.size();
                std::vector<int> sptr((size_t)ns + 1, 0);
                for (int q = 0; q < ns; q++)
                    sptr[q + 1] = sptr[q] + (ptr[job.spilled[q] + 1] - ptr[job.spilled[q]]);
                std::vector<int> sidx((size_t)sptr[ns]);
                std::vector<double> sval((size_t)sptr[ns]);
                for (int q = 0; q < ns; q++) {
                    const int p0 = ptr[job.spilled[q]], len = sptr[q + 1] - sptr[q];
                    std::copy(idx + p0, idx + p0 + len, sidx.begin() + sptr[q]);
                    std::copy(val + p0, val + p0 + len, sval.begin() + sptr[q]);
                }
                sp.nseg = ns;
                sp.A.nseg = ns;
                sp.A.nnz = sptr[ns];
                IPXGPU_TRY(upload(&sp.A.ptr, sptr, s));
                IPXGPU_TRY(upload(&sp.A.idx, sidx, s));
                IPXGPU_TRY(upload(&sp.A.val, sval, s));
                IPXGPU_TRY(upload(&sp.map, job.spilled, s));
                HostTiles st = build_tiles(sptr.data(), 0, ns, 0, 8 * kTileNnz);
                IPXGPU_TRY(upload_tiles(&sp.tiles, st, s));
                max_grid = std::max(max_grid, sp.tiles.ntiles);
                IPXGPU_CUDA(cudaStreamSynchronize(s));  // host vectors die here
            }
        }
        lap("upload band layouts");
        IPXGPU_TRY(dev_alloc(&c->t, (size_t)nloc));
        IPXGPU_TRY(dev_alloc(&c->xin, (size_t)m));
        IPXGPU_TRY(dev_alloc(&c->ybuf, (size_t)m + 1));
        IPXGPU_TRY(dev_alloc(&c->diag, (size_t)m));
        IPXGPU_TRY(dev_alloc(&c->W_own, (size_t)(nloc + m)));
        IPXGPU_TRY(dev_alloc(&c->st_dev, 1));
        IPXGPU_TRY(dev_alloc(&c->scalars, 8));
        IPXGPU_TRY(ensure_reduce(c, std::max(max_grid, c->num_sms * 8)));
        IPXGPU_CUDA(cudaHostAlloc((void**)&c->mirror_host, sizeof(HostMirror),
                                  cudaHostAllocMapped));
        std::memset(c->mirror_host, 0, sizeof(HostMirror));
        IPXGPU_CUDA(cudaHostGetDevicePointer((void**)&c->mirror_dev, c->mirror_host, 0));
        IPXGPU_CUDA(cudaStreamSynchronize(s));
        lap("work vectors");
    } catch (const std::bad_alloc&) {
        return fail(IPXGPU_ERR_OUT_OF_MEMORY, "host allocation failed while building layouts");
    }
    guard.c = nullptr;
    *out = c;
    return IPXGPU_OK;
}

int ipxgpu_get_layout(ipxgpu_ctx* c, int64_t out[8]) {
    if (!c || !out) return fail(IPXGPU_ERR_ARGUMENT, "null argument");
    if (c->group) return ipxgpu_get_layout(c->group->sub[0], out);
    int64_t ct = 0, rt = 0;
    for (const Panel& P : c->panels) {
        ct += P.col_tiles.ntiles;
        rt += P.row_tiles.ntiles;
    }
    out[0] = c->m; out[1] = c->n; out[2] = c->csc.nnz; out[3] = c->col_begin;
    out[4] = c->col_end; out[5] = (int64_t)c->panels.size(); out[6] = ct; out[7] = rt;
    return IPXGPU_OK;
}

int ipxgpu_get_tiling(ipxgpu_ctx* c, int64_t out[16]) {
    if (c && c->group) return ipxgpu_get_tiling(c->group->sub[0], out);
    if (!c || !out) return fail(IPXGPU_ERR_ARGUMENT, "null argument");
    const BandDev* ts[2] = {c->band1, c->band2};
    for (int k = 0; k < 2; k++) {
        int64_t* o = out + 8 * k;
        for (int j = 0; j < 8; j++) o[j] = 0;
        if (!ts[k]) continue;
        const BandPlan& P = ts[k]->plan;
        const Spill& sp = k == 0 ? c->spill1 : c->spill2;
        o[0] = sp.nseg > 0 ? 2 : 1;  // 2: banded, with spilled (dense) segments swept generically
        o[1] = P.VB; o[2] = P.SB; o[3] = P.NVB; o[4] = P.NSB;
        o[5] = P.K; o[6] = P.nparts; o[7] = P.nitems;
    }
    return IPXGPU_OK;
}

int ipxgpu_synchronize(ipxgpu_ctx* c) {
    IPXGPU_TRY(check_ctx(c));
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    return IPXGPU_OK;
}

int ipxgpu_launch_count(ipxgpu_ctx* c, int64_t* count) {
    if (!c || !count) return fail(IPXGPU_ERR_ARGUMENT, "null argument");
    if (c->group) {
        *count = 0;
        for (ipxgpu_ctx* s : c->group->sub) *count += s->launches;
        return IPXGPU_OK;
    }
    *count = c->launches;
    return IPXGPU_OK;
}

// ---- NCCL ----

int ipxgpu_comm_unique_id(char id[128]) {
    NcclApi* api = nccl_api();
    if (!api) return fail(IPXGPU_ERR_NCCL, "libnccl.so.2 not found");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId uid;
    ncclResult_t r = api->GetUniqueId(&uid);
    if (r != ncclSuccess) return fail(IPXGPU_ERR_NCCL, "ncclGetUniqueId failed");
    std::memcpy(id, &uid, 128);
    return IPXGPU_OK;
}

// Exchange buffer of a sharded context (plain cudaMalloc: IPC-exportable and peer-mappable).
static int ensure_xchg(ipxgpu_ctx* c) {
    if (c->nranks < 2 || c->nranks > 16) return fail(IPXGPU_ERR_STATE, "peer exchange needs 2..16 ranks");
    if (c->xchg) return IPXGPU_OK;
    c->xchg_mpad = ((size_t)c->m + 31) & ~(size_t)31;
    // [pull exchange: y[2][mpad] doubles, flags[nranks][SMs]] [push exchange:
    // records[2][nranks][mpad] of 16 bytes]
    size_t bytes = 2 * c->xchg_mpad * sizeof(double) +
                   (size_t)c->nranks * c->num_sms * sizeof(unsigned);
    bytes = (bytes + 255) & ~(size_t)255;
    c->xchg_ll_off = bytes;
    bytes += 2 * (size_t)c->nranks * c->xchg_mpad * 16;  // partial-product records
    bytes += 2 * c->xchg_mpad * 16;                       // final records (two-phase exchange)
    IPXGPU_CUDA(cudaMalloc(&c->xchg, bytes));
    IPXGPU_CUDA(cudaMemset(c->xchg, 0, bytes));
    c->xchg_gen = 0;
    IPXGPU_TRY(dev_alloc(&c->xchg_abort, 1));
    IPXGPU_CUDA(cudaMemset(c->xchg_abort, 0, sizeof(double)));
    return IPXGPU_OK;
}

int ipxgpu_peer_export(ipxgpu_ctx* c, char handle[64]) {
    IPXGPU_TRY(check_ctx(c));
    if (!handle) return fail(IPXGPU_ERR_ARGUMENT, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    IPXGPU_TRY(ensure_xchg(c));
    cudaIpcMemHandle_t h;
    IPXGPU_CUDA(cudaIpcGetMemHandle(&h, c->xchg));
    std::memcpy(handle, &h, 64);
    return IPXGPU_OK;
}

int ipxgpu_peer_import(ipxgpu_ctx* c, const char* handles) {
    IPXGPU_TRY(check_ctx(c));
    if (!handles) return fail(IPXGPU_ERR_ARGUMENT, "null argument");
    if (!c->xchg) return fail(IPXGPU_ERR_STATE, "call ipxgpu_peer_export first");
    for (int r = 0; r < c->nranks; r++) {
        if (r == c->rank) {
            c->peer_base[r] = c->xchg;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * 64, 64);
        IPXGPU_CUDA(cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    IPXGPU_TRY(dev_alloc(&c->peer_dev, (size_t)c->nranks));
    IPXGPU_CUDA(cudaMemcpy(c->peer_dev, c->peer_base, sizeof(void*) * c->nranks,
                           cudaMemcpyHostToDevice));
    c->peers_ready = true;
    return IPXGPU_OK;
}

int ipxgpu_comm_init(ipxgpu_ctx* c, const char id[128]) {
    IPXGPU_TRY(check_ctx(c));
    NcclApi* api = nccl_api();
    if (!api) return fail(IPXGPU_ERR_NCCL, "libnccl.so.2 not found");
    ncclUniqueId uid;
    std::memcpy(&uid, id, 128);
    ncclComm_t comm;
    ncclResult_t r = api->CommInitRank(&comm, c->nranks, uid, c->rank);
    if (r != ncclSuccess)
        return fail(IPXGPU_ERR_NCCL, std::string("ncclCommInitRank: ") +
                                         (api->GetErrorString ? api->GetErrorString(r) : "?"));
    c->nccl_comm = comm;
    return IPXGPU_OK;
}

// ---- NormalMatrix ----

int ipxgpu_normal_prepare(ipxgpu_ctx* c, const double* W) {
    IPXGPU_TRY(check_ctx(c));
    if (W) {
        IPXGPU_CUDA(cudaMemcpyAsync(c->W_own, W + c->col_begin, sizeof(double) * c->nloc,
                                    cudaMemcpyHostToDevice, c->stream));
        IPXGPU_CUDA(cudaMemcpyAsync(c->W_own + c->nloc, W + c->n, sizeof(double) * c->m,
                                    cudaMemcpyHostToDevice, c->stream));
        IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
        c->Wc = c->W_own;
        c->Ws = c->W_own + c->nloc;
    } else {
        c->Wc = nullptr;
        c->Ws = nullptr;
    }
    c->prepared = true;
    return IPXGPU_OK;
}

int ipxgpu_normal_prepare_dev(ipxgpu_ctx* c, const void* W_dev) {
    IPXGPU_TRY(check_ctx(c));
    const double* W = (const double*)W_dev;
    c->Wc = W ? W + c->col_begin : nullptr;
    c->Ws = W ? W + c->n : nullptr;
    c->prepared = true;
    return IPXGPU_OK;
}

int ipxgpu_normal_apply_dev(ipxgpu_ctx* c, const void* rhs_dev, void* lhs_dev) {
    IPXGPU_TRY(check_ctx(c));
    if (!c->prepared) return fail(IPXGPU_ERR_STATE, "normal matrix not prepared");
    return launch_normal_apply(c, (const double*)rhs_dev, (double*)lhs_dev, kApplyPlain,
                               kSlotNone, nullptr);
}

int ipxgpu_normal_apply(ipxgpu_ctx* c, const double* rhs, double* lhs, double* rhs_dot_lhs) {
    IPXGPU_TRY(check_ctx(c));
    if (!c->prepared) return fail(IPXGPU_ERR_STATE, "normal matrix not prepared");
    if (!rhs || !lhs) return fail(IPXGPU_ERR_ARGUMENT, "null vector");
    const size_t m = (size_t)c->m;
    IPXGPU_CUDA(cudaMemcpyAsync(c->xin, rhs, sizeof(double) * m, cudaMemcpyHostToDevice,
                                c->stream));
    IPXGPU_TRY(launch_normal_apply(c, c->xin, c->ybuf, kApplyPlain, kSlotNone, nullptr));
    IPXGPU_CUDA(cudaMemcpyAsync(lhs, c->ybuf, sizeof(double) * m, cudaMemcpyDeviceToHost,
                                c->stream));
    double dot = 0.0;
    if (rhs_dot_lhs && m > 0)
        IPXGPU_CUDA(cudaMemcpyAsync(&dot, c->ybuf + m, sizeof(double), cudaMemcpyDeviceToHost,
                                    c->stream));
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    if (rhs_dot_lhs) *rhs_dot_lhs = dot;
    return IPXGPU_OK;
}

// ---- DiagonalPrecond ----

// Device pointers of the weights a diagonal build uses: the prepared ones, a staged upload of
// W, or none (W = 1 on structurals, 0 on slacks).
static int diag_weights(ipxgpu_ctx* c, const double* W, int use_prepared, const double** Wc,
                        const double** Ws) {
    if (use_prepared) {
        if (!c->prepared) return fail(IPXGPU_ERR_STATE, "normal matrix not prepared");
        *Wc = c->Wc;
        *Ws = c->Ws;
    } else if (W) {
        // Stage into the tail of the n-vector scratch so the prepared weights
        // of the normal matrix stay intact.
        IPXGPU_TRY(ensure_nvecs(c));
        double* stage = c->nvec[0];
        IPXGPU_CUDA(cudaMemcpyAsync(stage, W + c->col_begin, sizeof(double) * c->nloc,
                                    cudaMemcpyHostToDevice, c->stream));
        IPXGPU_CUDA(cudaMemcpyAsync(stage + c->nloc, W + c->n, sizeof(double) * c->m,
                                    cudaMemcpyHostToDevice, c->stream));
        *Wc = stage;
        *Ws = stage + c->nloc;
    } else {
        *Wc = nullptr;
        *Ws = nullptr;
    }
    return IPXGPU_OK;
}

static void free_smw(ipxgpu_ctx* c) {
    SmwDev& S = c->smw;
    dev_free(S.colptr); dev_free(S.rowidx); dev_free(S.colval);
    dev_free(S.rowptr); dev_free(S.colidx); dev_free(S.rowval);
    dev_free(S.chunk_col); dev_free(S.chunk_p0); dev_free(S.col_chunk0);
    dev_free(S.partials); dev_free(S.L); dev_free(S.z);
    S = SmwDev();
    c->smw_active = false;
}

int ipxgpu_diag_factorize(ipxgpu_ctx* c, const double* W, int use_prepared) {
    IPXGPU_TRY(check_ctx(c));
    const double *Wc, *Ws;
    IPXGPU_TRY(diag_weights(c, W, use_prepared, &Wc, &Ws));
    if (c->panels.empty())
        IPXGPU_CUDA(cudaMemsetAsync(c->diag, 0, sizeof(double) * std::max<int64_t>(1, c->m),
                                    c->stream));
    IPXGPU_TRY(launch_diag_build(c, Wc, Ws));
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    c->smw_active = false;
    return IPXGPU_OK;
}

__global__ void __launch_bounds__(kBlock)
zero_entries_kernel(int count, const int* __restrict__ where, double* x) {
    const int k = blockIdx.x * kBlock + threadIdx.x;
    if (k < count) x[where[k]] = 0.0;
}

int ipxgpu_diag_factorize_masked(ipxgpu_ctx* c, const double* W, int use_prepared, int64_t nd,
                                 const int64_t* dense_cols) {
    IPXGPU_TRY(check_ctx(c));
    if (nd < 0 || (nd > 0 && !dense_cols)) return fail(IPXGPU_ERR_ARGUMENT, "dense column list");
    const double *Wc, *Ws;
    IPXGPU_TRY(diag_weights(c, W, use_prepared, &Wc, &Ws));
    // The shard's structural weights with the dense columns' entries zeroed: those columns
    // never enter the sum (reference src/diagonal_precond.cc:28-36).
    std::vector<int> local;
    for (int64_t k = 0; k < nd; k++) {
        const int64_t j = dense_cols[k];
        if (j < 0 || j >= c->n) return fail(IPXGPU_ERR_ARGUMENT, "dense column out of range");
        if (j >= c->col_begin && j < c->col_end) local.push_back((int)(j - c->col_begin));
    }
    if (c->nloc > 0) {
        if (!c->W_mask) IPXGPU_TRY(dev_alloc(&c->W_mask, (size_t)c->nloc));
        if (Wc)
            IPXGPU_CUDA(cudaMemcpyAsync(c->W_mask, Wc, sizeof(double) * c->nloc,
                                        cudaMemcpyDeviceToDevice, c->stream));
        else
            fill_kernel<<<grid_for(c, c->nloc), kBlock, 0, c->stream>>>(c->nloc, c->W_mask, 1.0);
        if (!local.empty()) {
            int* where = nullptr;
            IPXGPU_TRY(upload(&where, local, c->stream));
            zero_entries_kernel<<<((int)local.size() + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(
                (int)local.size(), where, c->W_mask);
            c->launches++;
            IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
            dev_free(where);
        }
    }
    if (c->panels.empty())
        IPXGPU_CUDA(cudaMemsetAsync(c->diag, 0, sizeof(double) * std::max<int64_t>(1, c->m),
                                    c->stream));
    IPXGPU_TRY(launch_diag_build(c, c->W_mask, Ws));
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    c->smw_active = false;  // until ipxgpu_smw_load installs the factor that belongs to this E
    return IPXGPU_OK;
}

int ipxgpu_smw_clear(ipxgpu_ctx* c) {
    IPXGPU_TRY(check_ctx(c));
    c->smw_active = false;
    return IPXGPU_OK;
}

int ipxgpu_smw_load(ipxgpu_ctx* c, int64_t nd, const int64_t* Adp, const int64_t* Adi,
                    const double* Adx, const double* L) {
    IPXGPU_TRY(check_ctx(c));
    if (nd <= 0 || nd > kSmwMaxCols || !Adp || !L)
        return fail(IPXGPU_ERR_ARGUMENT, "ipxgpu_smw_load: 1 <= nd <= 1024 dense columns");
    const int m = (int)c->m;
    const int64_t nnz = Adp[nd];
    if (nnz > 0 && (!Adi || !Adx)) return fail(IPXGPU_ERR_ARGUMENT, "ipxgpu_smw_load: null entries");
    if (nnz >= (int64_t)INT32_MAX) return fail(IPXGPU_ERR_UNSUPPORTED, "dense columns too large");
    if (!c->diag_ready) return fail(IPXGPU_ERR_STATE, "diagonal not factorized");
    free_smw(c);
    SmwDev& S = c->smw;
    S.nd = (int)nd;
    S.m = m;
    // by column, by row (columns ascending inside a row), and the gather chunks
    std::vector<int> colptr((size_t)nd + 1), rowidx((size_t)nnz), rowptr((size_t)m + 1, 0),
        colidx((size_t)nnz), chunk_col, chunk_p0, col_chunk0((size_t)nd + 1);
    std::vector<double> rowval((size_t)nnz);
    for (int64_t k = 0; k <= nd; k++) colptr[k] = (int)Adp[k];
    for (int64_t p = 0; p < nnz; p++) {
        if (Adi[p] < 0 || Adi[p] >= m) return fail(IPXGPU_ERR_ARGUMENT, "dense column: row index");
        rowidx[p] = (int)Adi[p];
        rowptr[Adi[p] + 1]++;
    }
    for (int i = 0; i < m; i++) rowptr[i + 1] += rowptr[i];
    {
        std::vector<int> next(rowptr.begin(), rowptr.end() - 1);
        for (int64_t k = 0; k < nd; k++)
            for (int64_t p = Adp[k]; p < Adp[k + 1]; p++) {
                const int q = next[Adi[p]]++;
                colidx[q] = (int)k;
                rowval[q] = Adx[p];
            }
    }
    for (int64_t k = 0; k < nd; k++) {
        col_chunk0[k] = (int)chunk_col.size();
        for (int64_t p = Adp[k]; p < Adp[k + 1]; p += kSmwChunk) {
            chunk_col.push_back((int)k);
            chunk_p0.push_back((int)p);
        }
        if (Adp[k] == Adp[k + 1]) {  // keep one (empty) chunk per column
            chunk_col.push_back((int)k);
            chunk_p0.push_back((int)Adp[k]);
        }
    }
    col_chunk0[nd] = (int)chunk_col.size();
    chunk_p0.push_back((int)nnz);
    S.nchunks = (int)chunk_col.size();
    std::vector<int> chunk_end((size_t)S.nchunks);
    for (int cidx = 0; cidx < S.nchunks; cidx++) {
        const int k = chunk_col[cidx];
        chunk_end[cidx] = std::min(colptr[k + 1], chunk_p0[cidx] + kSmwChunk);
    }
    // The kernel reads a chunk's range as [chunk_p0[c], chunk_p0[c+1]): columns are contiguous
    // in Adp and chunks are in entry order, so chunk c+1 begins where chunk c ends.
    for (int cidx = 0; cidx + 1 < S.nchunks; cidx++)
        if (chunk_end[cidx] != chunk_p0[cidx + 1])
            return fail(IPXGPU_ERR_ARGUMENT, "ipxgpu_smw_load: Adp not contiguous");
    std::vector<double> colval(Adx, Adx + nnz);
    IPXGPU_TRY(upload(&S.colptr, colptr, c->stream));
    IPXGPU_TRY(upload(&S.rowidx, rowidx, c->stream));
    IPXGPU_TRY(upload(&S.colval, colval, c->stream));
    IPXGPU_TRY(upload(&S.rowptr, rowptr, c->stream));
    IPXGPU_TRY(upload(&S.colidx, colidx, c->stream));
    IPXGPU_TRY(upload(&S.rowval, rowval, c->stream));
    IPXGPU_TRY(upload(&S.chunk_col, chunk_col, c->stream));
    IPXGPU_TRY(upload(&S.chunk_p0, chunk_p0, c->stream));
    IPXGPU_TRY(upload(&S.col_chunk0, col_chunk0, c->stream));
    IPXGPU_TRY(dev_alloc(&S.partials, 2 * (size_t)S.nchunks));
    IPXGPU_TRY(dev_alloc(&S.z, 2 * (size_t)kSmwMaxCols));
    IPXGPU_TRY(dev_alloc(&S.L, (size_t)nd * nd));
    IPXGPU_CUDA(cudaMemcpyAsync(S.L, L, sizeof(double) * nd * nd, cudaMemcpyHostToDevice,
                                c->stream));
    S.warp_rows = nnz > 16 * (int64_t)std::max(1, m) ? 1 : 0;
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    c->smw_active = true;
    return IPXGPU_OK;
}

int ipxgpu_diag_get(ipxgpu_ctx* c, double* diag) {
    IPXGPU_TRY(check_ctx(c));
    if (!c->diag_ready) return fail(IPXGPU_ERR_STATE, "diagonal not factorized");
    IPXGPU_CUDA(cudaMemcpyAsync(diag, c->diag, sizeof(double) * c->m, cudaMemcpyDeviceToHost,
                                c->stream));
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    return IPXGPU_OK;
}

int ipxgpu_diag_set(ipxgpu_ctx* c, const double* diag) {
    IPXGPU_TRY(check_ctx(c));
    IPXGPU_CUDA(cudaMemcpyAsync(c->diag, diag, sizeof(double) * c->m, cudaMemcpyHostToDevice,
                                c->stream));
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    c->diag_ready = true;
    return IPXGPU_OK;
}

int ipxgpu_diag_apply(ipxgpu_ctx* c, const double* rhs, double* lhs, double* rhs_dot_lhs) {
    IPXGPU_TRY(check_ctx(c));
    if (!c->diag_ready) return fail(IPXGPU_ERR_STATE, "diagonal not factorized");
    const int m = (int)c->m;
    const int grid = grid_for(c, m);
    IPXGPU_TRY(ensure_reduce(c, grid));
    IPXGPU_CUDA(cudaMemcpyAsync(c->xin, rhs, sizeof(double) * m, cudaMemcpyHostToDevice,
                                c->stream));
    if (c->smw_active) {
        const int sgrid = smw_grid(c);
        IPXGPU_TRY(ensure_reduce(c, sgrid));
        IPXGPU_TRY(launch_smw_pass(c, c->xin, nullptr, kSmwFirst, nullptr));
        smw_apply_finish_kernel<<<sgrid, kBlock, 0, c->stream>>>(c->smw, c->diag, c->xin, c->ybuf,
                                                                 c->red, c->ybuf + m);
    } else {
        diag_apply_kernel<<<grid, kBlock, 0, c->stream>>>(m, c->diag, c->xin, c->ybuf, c->red,
                                                          c->ybuf + m);
    }
    c->launches++;
    IPXGPU_CUDA(cudaGetLastError());
    IPXGPU_CUDA(cudaMemcpyAsync(lhs, c->ybuf, sizeof(double) * m, cudaMemcpyDeviceToHost,
                                c->stream));
    double dot = 0.0;
    IPXGPU_CUDA(cudaMemcpyAsync(&dot, c->ybuf + m, sizeof(double), cudaMemcpyDeviceToHost,
                                c->stream));
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    if (rhs_dot_lhs) *rhs_dot_lhs = dot;
    return IPXGPU_OK;
}

// ---- ConjugateResiduals ----

static int cr_host_entry(ipxgpu_ctx* c, int op, bool precond, const double* rhs, double tol,
                         const double* resscale, int64_t maxiter, double* lhs,
                         ipxgpu_cr_result* result, ipxgpu_interrupt_fn interrupt, void* user,
                         double* hist, int64_t hist_cap) {
    IPXGPU_TRY(check_ctx(c));
    if (!rhs || !lhs) return fail(IPXGPU_ERR_ARGUMENT, "null vector");
    if (op == 0 && !c->prepared) return fail(IPXGPU_ERR_STATE, "normal matrix not prepared");
    if (op == 1 && !split_ready(c)) return fail(IPXGPU_ERR_STATE, "split operator not prepared");
    if (op != 0 && op != 1) return fail(IPXGPU_ERR_ARGUMENT, "unknown operator");
    if (precond && !c->diag_ready) return fail(IPXGPU_ERR_STATE, "diagonal not factorized");
    const auto t0 = std::chrono::steady_clock::now();
    if (!hist) hist_cap = 0;
    IPXGPU_TRY(ensure_cr_buffers(c, hist_cap));
    const size_t m = (size_t)c->m;
    // Infnorm(lhs) == 0 saves a matrix-vector product
    // (reference src/conjugate_residuals.cc:33, :118).
    bool zero_start = true;
    for (size_t i = 0; i < m; i++)
        if (std::fabs(lhs[i]) > 0.0 || lhs[i] != lhs[i]) { zero_start = false; break; }
    IPXGPU_CUDA(cudaMemcpyAsync(c->v_rhs, rhs, sizeof(double) * m, cudaMemcpyHostToDevice,
                                c->stream));
    if (!zero_start)
        IPXGPU_CUDA(cudaMemcpyAsync(c->v_y, lhs, sizeof(double) * m, cudaMemcpyHostToDevice,
                                    c->stream));
    if (resscale)
        IPXGPU_CUDA(cudaMemcpyAsync(c->v_resscale, resscale, sizeof(double) * m,
                                    cudaMemcpyHostToDevice, c->stream));
    if (hist_cap > 0)
        IPXGPU_CUDA(cudaMemsetAsync(c->v_hist, 0xff, sizeof(double) * hist_cap, c->stream));
    IPXGPU_TRY(run_cr(c, op, precond, zero_start, resscale != nullptr, tol, maxiter, result,
                      interrupt, user, hist_cap));
    IPXGPU_CUDA(cudaMemcpy(lhs, c->v_y, sizeof(double) * m, cudaMemcpyDeviceToHost));
    if (hist_cap > 0)
        IPXGPU_CUDA(cudaMemcpy(hist, c->v_hist, sizeof(double) * hist_cap,
                               cudaMemcpyDeviceToHost));
    if (result)
        result->time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return op == 1 ? check_tri(c) : IPXGPU_OK;
}

int ipxgpu_pcr_solve(ipxgpu_ctx* c, const double* rhs, double tol, const double* resscale,
                     int64_t maxiter, double* lhs, ipxgpu_cr_result* result,
                     ipxgpu_interrupt_fn interrupt, void* user, double* resnorm_hist,
                     int64_t hist_cap) {
    return cr_host_entry(c, 0, true, rhs, tol, resscale, maxiter, lhs, result, interrupt, user,
                         resnorm_hist, hist_cap);
}

int ipxgpu_pcr_solve_dev(ipxgpu_ctx* c, const void* rhs_dev, double tol, const void* resscale_dev,
                         int64_t maxiter, void* lhs_dev, int zero_start,
                         ipxgpu_cr_result* result) {
    IPXGPU_TRY(check_ctx(c));
    if (!rhs_dev || !lhs_dev) return fail(IPXGPU_ERR_ARGUMENT, "null vector");
    if (!c->prepared) return fail(IPXGPU_ERR_STATE, "normal matrix not prepared");
    if (!c->diag_ready) return fail(IPXGPU_ERR_STATE, "diagonal not factorized");
    const auto t0 = std::chrono::steady_clock::now();
    IPXGPU_TRY(ensure_cr_buffers(c, 0));
    const size_t bytes = sizeof(double) * (size_t)c->m;
    cudaStream_t s = c->stream;
    IPXGPU_CUDA(cudaMemcpyAsync(c->v_rhs, rhs_dev, bytes, cudaMemcpyDeviceToDevice, s));
    if (!zero_start)
        IPXGPU_CUDA(cudaMemcpyAsync(c->v_y, lhs_dev, bytes, cudaMemcpyDeviceToDevice, s));
    if (resscale_dev)
        IPXGPU_CUDA(cudaMemcpyAsync(c->v_resscale, resscale_dev, bytes,
                                    cudaMemcpyDeviceToDevice, s));
    IPXGPU_TRY(run_cr(c, 0, true, zero_start != 0, resscale_dev != nullptr, tol, maxiter, result,
                      nullptr, nullptr, 0));
    IPXGPU_CUDA(cudaMemcpyAsync(lhs_dev, c->v_y, bytes, cudaMemcpyDeviceToDevice, s));
    IPXGPU_CUDA(cudaStreamSynchronize(s));
    if (result)
        result->time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return IPXGPU_OK;
}

int ipxgpu_cr_solve(ipxgpu_ctx* c, int op, const double* rhs, double tol, const double* resscale,
                    int64_t maxiter, double* lhs, ipxgpu_cr_result* result,
                    ipxgpu_interrupt_fn interrupt, void* user, double* resnorm_hist,
                    int64_t hist_cap) {
    return cr_host_entry(c, op, false, rhs, tol, resscale, maxiter, lhs, result, interrupt, user,
                         resnorm_hist, hist_cap);
}

// ---- KKTSolverDiag ----

// Largest of a scalar over the ranks (NCCL max; nranks == 1: nothing to do).
static int allreduce_max(ipxgpu_ctx* c, double* buf, size_t count) {
    if (c->nranks == 1) return IPXGPU_OK;
    if (!c->nccl_comm)
        return fail(IPXGPU_ERR_STATE, "nranks > 1 but ipxgpu_comm_init was not called");
    NcclApi* api = nccl_api();
    ncclResult_t r = api->AllReduce(buf, buf, count, ncclDouble, ncclMax,
                                    (ncclComm_t)c->nccl_comm, c->stream);
    if (r != ncclSuccess)
        return fail(IPXGPU_ERR_NCCL, std::string("ncclAllReduce(max): ") +
                                         (api->GetErrorString ? api->GetErrorString(r) : "?"));
    return IPXGPU_OK;
}

// The KKTSolverDiag entry points work in the shard's LOCAL layout [nloc structural columns of
// the shard | m slack columns]: with one rank that is the whole (n+m)-vector; with column
// shards every rank holds its columns' part of the iterate, of W, of a and of x plus a replica
// of the slack part, and the three m-vectors that sum over columns (diagonal, right-hand side,
// slack part of x) are allreduced.
static int upload_local(ipxgpu_ctx* c, double* dst, const double* src_full) {
    if (c->nloc > 0)
        IPXGPU_CUDA(cudaMemcpyAsync(dst, src_full + c->col_begin, sizeof(double) * c->nloc,
                                    cudaMemcpyHostToDevice, c->stream));
    if (c->m > 0)
        IPXGPU_CUDA(cudaMemcpyAsync(dst + c->nloc, src_full + c->n, sizeof(double) * c->m,
                                    cudaMemcpyHostToDevice, c->stream));
    return IPXGPU_OK;
}

int ipxgpu_kktdiag_factorize(ipxgpu_ctx* c, const double* xl, const double* xu, const double* zl,
                             const double* zu, double mu, double* W_out, double* resscale_out) {
    if (c && c->group)
        return group_kktdiag_factorize(c, xl, xu, zl, zu, mu, W_out, resscale_out);
    IPXGPU_TRY(check_ctx(c));
    const long long nl = (long long)c->nloc + c->m;  // local layout
    if (!c->W_full) IPXGPU_TRY(dev_alloc(&c->W_full, (size_t)(c->n + c->m)));
    if (!c->resscale_kkt) IPXGPU_TRY(dev_alloc(&c->resscale_kkt, (size_t)c->m));
    const int grid = grid_for(c, nl);
    IPXGPU_TRY(ensure_reduce(c, grid));
    c->kkt_factorized = false;
    if (xl) {
        if (!xu || !zl || !zu) return fail(IPXGPU_ERR_ARGUMENT, "incomplete iterate");
        IPXGPU_TRY(ensure_nvecs(c));
        // The four iterate vectors stream through persistent scratch vectors (no
        // allocation per call: cudaMalloc/cudaFree cost up to hundreds of ms).
        double* d_xl = c->nvec[0];
        double* d_xu = c->nvec[1];
        double* d_zl = c->nvec[2];
        double* d_zu = c->nvec[3];
        cudaStream_t s = c->stream;
        IPXGPU_TRY(upload_local(c, d_xl, xl));
        IPXGPU_TRY(upload_local(c, d_xu, xu));
        IPXGPU_TRY(upload_local(c, d_zl, zl));
        IPXGPU_TRY(upload_local(c, d_zu, zu));
        kkt_weights_kernel<<<grid, kBlock, 0, s>>>(nl, d_xl, d_xu, d_zl, d_zu, c->W_full, c->red,
                                                   c->scalars);
        // regval is the smallest nonzero g over ALL columns (reference :34-41)
        IPXGPU_TRY(allreduce_max(c, c->scalars, 1));
        kkt_weights_fix_kernel<<<grid, kBlock, 0, s>>>(nl, c->nloc, c->W_full, c->resscale_kkt, mu,
                                                       c->scalars);
        c->launches += 2;
        IPXGPU_CUDA(cudaStreamSynchronize(s));
    } else {
        fill_kernel<<<grid, kBlock, 0, c->stream>>>(nl, c->W_full, 1.0);
        fill_kernel<<<grid_for(c, c->m), kBlock, 0, c->stream>>>(c->m, c->resscale_kkt, 1.0);
        c->launches += 2;
    }
    IPXGPU_CUDA(cudaGetLastError());
    c->Wc = c->W_full;
    c->Ws = c->W_full + c->nloc;
    c->prepared = true;
    IPXGPU_TRY(launch_diag_build(c, c->Wc, c->Ws));
    c->smw_active = false;  // a dense-column part belongs to the diagonal it was built for
    if (W_out) {
        // this rank's columns; the slack part once (identical on every rank)
        if (c->nloc > 0)
            IPXGPU_CUDA(cudaMemcpyAsync(W_out + c->col_begin, c->W_full, sizeof(double) * c->nloc,
                                        cudaMemcpyDeviceToHost, c->stream));
        if (c->rank == 0 && c->m > 0)
            IPXGPU_CUDA(cudaMemcpyAsync(W_out + c->n, c->W_full + c->nloc, sizeof(double) * c->m,
                                        cudaMemcpyDeviceToHost, c->stream));
    }
    if (resscale_out && c->rank == 0)
        IPXGPU_CUDA(cudaMemcpyAsync(resscale_out, c->resscale_kkt, sizeof(double) * c->m,
                                    cudaMemcpyDeviceToHost, c->stream));
    IPXGPU_CUDA(cudaStreamSynchronize(c->stream));
    c->kkt_factorized = true;
    return IPXGPU_OK;
}

int ipxgpu_kktdiag_solve(ipxgpu_ctx* c, const double* a, const double* b, double tol,
                         int64_t maxiter, double* x, double* y, ipxgpu_cr_result* result,
                         ipxgpu_interrupt_fn interrupt, void* user) {
    if (c && c->group)
        return group_kktdiag_solve(c, a, b, tol, maxiter, x, y, result, interrupt, user);
    IPXGPU_TRY(check_ctx(c));
    if (!c->kkt_factorized) return fail(IPXGPU_ERR_STATE, "kktdiag not factorized");
    if (!a || !b || !x || !y) return fail(IPXGPU_ERR_ARGUMENT, "null vector");
    const auto t0 = std::chrono::steady_clock::now();
    const long long nloc = c->nloc, m = c->m;
    const bool lead = c->rank == 0;
    IPXGPU_TRY(ensure_nvecs(c));
    IPXGPU_TRY(ensure_cr_buffers(c, 0));
    cudaStream_t s = c->stream;
    double* d_a = c->nvec[0];
    double* d_u = c->nvec[1];  // W.*a, later x
    double* d_b = c->xin;
    IPXGPU_TRY(upload_local(c, d_a, a));
    IPXGPU_CUDA(cudaMemcpyAsync(d_b, b, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    // rhs = -b + AI*(W.*a)  (reference src/kkt_solver_diag.cc:90-92); the slack columns' term
    // and -b enter once (rank 0), the structural columns' terms are summed over the ranks
    mul_kernel<<<grid_for(c, nloc), kBlock, 0, s>>>(nloc, c->W_full, d_a, d_u);
    if (lead)
        mul_sub_kernel<<<grid_for(c, m), kBlock, 0, s>>>(m, c->W_full + nloc, d_a + nloc, d_b,
                                                         c->ybuf);
    else
        IPXGPU_CUDA(cudaMemsetAsync(c->ybuf, 0, sizeof(double) * m, s));
    c->launches += 2;
    if (c->panels.empty())
        IPXGPU_CUDA(cudaMemcpyAsync(c->v_rhs, c->ybuf, sizeof(double) * m,
                                    cudaMemcpyDeviceToDevice, s));
    for (size_t k = 0; k < c->panels.size(); k++) {
        const Panel& P = c->panels[k];
        OpRowAffine op{d_u, c->ybuf, c->v_rhs, 1.0, k == 0};
        IPXGPU_TRY(launch_sweep(c, op, P.row_tiles, P.csr, nullptr));
    }
    IPXGPU_TRY(allreduce_sum(c, c->v_rhs, (size_t)m));
    // y = 0; PCR with resscale (reference :95-99)
    IPXGPU_CUDA(cudaMemcpyAsync(c->v_resscale, c->resscale_kkt, sizeof(double) * m,
                                cudaMemcpyDeviceToDevice, s));
    IPXGPU_TRY(run_cr(c, 0, true, true, true, tol, maxiter, result, interrupt, user, 0));
    // Recovery (reference :108-117): x[j] = W[j]*(a[j] - A[:,j]'y);
    // x[n+i] = b[i] - sum_j x[j] a_ij (the sum over all ranks' columns).
    for (size_t k = 0; k < c->panels.size(); k++) {
        const Panel& P = c->panels[k];
        OpColRecover op{c->v_y, c->W_full, d_a, d_u};
        IPXGPU_TRY(launch_sweep(c, op, P.col_tiles, c->csc, nullptr));
    }
    double* d_xs = d_u + nloc;  // slack part of x
    if (c->nranks == 1) {
        if (c->panels.empty())
            IPXGPU_CUDA(cudaMemcpyAsync(d_xs, d_b, sizeof(double) * m, cudaMemcpyDeviceToDevice, s));
        for (size_t k = 0; k < c->panels.size(); k++) {
            const Panel& P = c->panels[k];
            OpRowAffine op{d_u, d_b, d_xs, -1.0, k == 0};
            IPXGPU_TRY(launch_sweep(c, op, P.row_tiles, P.csr, nullptr));
        }
    } else {
        // partial sums A_g x_g (from zero), summed over the ranks, then b - sum
        IPXGPU_CUDA(cudaMemsetAsync(c->ybuf, 0, sizeof(double) * m, s));
        if (c->panels.empty())
            IPXGPU_CUDA(cudaMemsetAsync(d_xs, 0, sizeof(double) * m, s));
        for (size_t k = 0; k < c->panels.size(); k++) {
            const Panel& P = c->panels[k];
            OpRowAffine op{d_u, c->ybuf, d_xs, 1.0, k == 0};
            IPXGPU_TRY(launch_sweep(c, op, P.row_tiles, P.csr, nullptr));
        }
        IPXGPU_TRY(allreduce_sum(c, d_xs, (size_t)m));
        sub_from_kernel<<<grid_for(c, m), kBlock, 0, s>>>(m, d_b, d_xs);
        c->launches++;
    }
    if (nloc > 0)
        IPXGPU_CUDA(cudaMemcpyAsync(x + c->col_begin, d_u, sizeof(double) * nloc,
                                    cudaMemcpyDeviceToHost, s));
    if (lead) {
        IPXGPU_CUDA(cudaMemcpyAsync(x + c->n, d_xs, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
        IPXGPU_CUDA(cudaMemcpyAsync(y, c->v_y, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
    }
    IPXGPU_CUDA(cudaStreamSynchronize(s));
    if (result)
        result->time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return IPXGPU_OK;
}


}  // extern "C"

// ---- one process, several GPUs (IPXGPU_NGPUS of the drop-in build) ----
//
// A group is G ordinary sharded contexts, one per device, each driven from its own host thread
// for the duration of a call (the calls block: the persistent CR kernels of the ranks meet once
// per iteration over NVLink). NCCL communicator per rank (ncclCommInitRank from G threads),
// exchange buffers mapped with cudaDeviceEnablePeerAccess instead of IPC handles. The group's
// handle is an ipxgpu_ctx without device data; it accepts the entry points KKTSolverDiag needs.
namespace {

// Runs fn(rank) on one thread per rank; returns the first failure (its message becomes this
// thread's last error).
template <class F>
int group_run(ipxgpu_group* g, F fn) {
    const int G = (int)g->sub.size();
    std::vector<int> rc((size_t)G, IPXGPU_OK);
    std::vector<std::string> msg((size_t)G);
    std::vector<std::thread> th;
    for (int r = 0; r < G; r++)
        th.emplace_back([&, r] {
            try {
                rc[r] = fn(r);
            } catch (const std::bad_alloc&) {
                rc[r] = fail(IPXGPU_ERR_OUT_OF_MEMORY, "host allocation failed");
            } catch (const std::exception& e) {
                rc[r] = fail(IPXGPU_ERR_STATE, e.what());
            }
            if (rc[r] != IPXGPU_OK) msg[r] = ipxgpu_last_error();
        });
    for (std::thread& t : th) t.join();
    for (int r = 0; r < G; r++)
        if (rc[r] != IPXGPU_OK) return fail(rc[r], "rank " + std::to_string(r) + ": " + msg[r]);
    return IPXGPU_OK;
}

}  // namespace

extern "C" {

static void destroy_group(ipxgpu_ctx* c) {
    ipxgpu_group* g = c->group;
    for (ipxgpu_ctx* s : g->sub) ipxgpu_destroy(s);
    delete g;
    c->group = nullptr;
    delete c;
}

int ipxgpu_create_group(ipxgpu_ctx** out, int64_t m, int64_t n, const int64_t* AIp,
                        const int64_t* AIi, const double* AIx, const ipxgpu_options* opt_in,
                        int32_t ngpus, const int32_t* devices) {
    if (!out) return fail(IPXGPU_ERR_ARGUMENT, "null argument");
    *out = nullptr;
    if (ngpus < 2 || ngpus > 16) return fail(IPXGPU_ERR_ARGUMENT, "a group has 2..16 GPUs");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < ngpus)
        return fail(IPXGPU_ERR_CUDA, "fewer CUDA devices than the group asks for");
    ipxgpu_options base;
    ipxgpu_default_options(&base);
    if (opt_in) base = *opt_in;
    if (base.stream) return fail(IPXGPU_ERR_ARGUMENT, "a group owns its streams");
    std::vector<int> dev((size_t)ngpus);
    for (int r = 0; r < ngpus; r++) {
        dev[r] = devices ? devices[r] : r;
        if (dev[r] < 0 || dev[r] >= ndev) return fail(IPXGPU_ERR_ARGUMENT, "device ordinal");
    }
    std::vector<int64_t> bounds((size_t)ngpus + 1);
    IPXGPU_TRY(ipxgpu_partition_columns(n, AIp, ngpus, bounds.data()));
    char uid[128];
    IPXGPU_TRY(ipxgpu_comm_unique_id(uid));
    ipxgpu_group* g = new ipxgpu_group;
    g->sub.assign((size_t)ngpus, nullptr);
    int rc = group_run(g, [&](int r) -> int {
        ipxgpu_options o = base;
        o.device = dev[r];
        o.rank = r;
        o.nranks = ngpus;
        o.col_begin = bounds[r];
        o.col_end = bounds[r + 1];
        IPXGPU_TRY(ipxgpu_create(&g->sub[r], m, n, AIp, AIi, AIx, &o));
        IPXGPU_TRY(ipxgpu_comm_init(g->sub[r], uid));  // all ranks meet inside NCCL
        for (int q = 0; q < ngpus; q++)
            if (q != r) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(dev[q], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess)
                    return fail(IPXGPU_ERR_CUDA, std::string("peer access: ") + cudaGetErrorString(e));
            }
        return ensure_xchg(g->sub[r]);
    });
    if (rc == IPXGPU_OK)
        rc = group_run(g, [&](int r) -> int {
            ipxgpu_ctx* c = g->sub[r];
            IPXGPU_TRY(check_ctx(c));
            for (int q = 0; q < ngpus; q++) c->peer_base[q] = g->sub[q]->xchg;
            c->peers_direct = true;
            IPXGPU_TRY(dev_alloc(&c->peer_dev, (size_t)ngpus));
            IPXGPU_CUDA(cudaMemcpy(c->peer_dev, c->peer_base, sizeof(void*) * ngpus,
                                   cudaMemcpyHostToDevice));
            c->peers_ready = true;
            return IPXGPU_OK;
        });
    if (rc != IPXGPU_OK) {
        const std::string why = ipxgpu_last_error();
        for (ipxgpu_ctx* s : g->sub) ipxgpu_destroy(s);
        delete g;
        return fail(rc, why);
    }
    ipxgpu_ctx* lead = new ipxgpu_ctx;
    lead->group = g;
    lead->m = m;
    lead->n = n;
    lead->nranks = ngpus;
    lead->device = dev[0];
    *out = lead;
    return IPXGPU_OK;
}

static int group_kktdiag_factorize(ipxgpu_ctx* c, const double* xl, const double* xu,
                                   const double* zl, const double* zu, double mu, double* W_out,
                                   double* resscale_out) {
    ipxgpu_group* g = c->group;
    return group_run(g, [&](int r) {
        return ipxgpu_kktdiag_factorize(g->sub[r], xl, xu, zl, zu, mu, W_out, resscale_out);
    });
}

namespace {
// The caller's interrupt callback is polled by rank 0 only; the other ranks follow its verdict,
// so that all of them stop (the ranks of a solve wait for each other once per iteration).
struct GroupInterrupt {
    ipxgpu_interrupt_fn fn;
    void* user;
    std::atomic<int64_t> verdict{0};
};
int64_t group_interrupt_lead(void* p) {
    GroupInterrupt* gi = static_cast<GroupInterrupt*>(p);
    int64_t v = gi->verdict.load();
    if (v == 0 && gi->fn) {
        v = gi->fn(gi->user);
        if (v != 0) gi->verdict.store(v);
    }
    return v;
}
int64_t group_interrupt_follow(void* p) { return static_cast<GroupInterrupt*>(p)->verdict.load(); }
}  // namespace

static int group_kktdiag_solve(ipxgpu_ctx* c, const double* a, const double* b, double tol,
                               int64_t maxiter, double* x, double* y, ipxgpu_cr_result* result,
                               ipxgpu_interrupt_fn interrupt, void* user) {
    ipxgpu_group* g = c->group;
    const int G = (int)g->sub.size();
    std::vector<ipxgpu_cr_result> res((size_t)G);
    GroupInterrupt gi;
    gi.fn = interrupt;
    gi.user = user;
    const int rc = group_run(g, [&](int r) {
        return ipxgpu_kktdiag_solve(g->sub[r], a, b, tol, maxiter, x, y, &res[r],
                                    r == 0 ? group_interrupt_lead : group_interrupt_follow, &gi);
    });
    const int64_t verdict = gi.verdict.load();
    if (result) {
        *result = res[0];
        if (verdict != 0) result->errflag = verdict;
    }
    // a rank that stopped on the caller's request leaves its peers waiting for an exchange
    // that never comes; they give up with an error that is the interrupt's, not a failure
    if (rc != IPXGPU_OK && verdict != 0) return IPXGPU_OK;
    return rc;
}

// ---- host-only layout check ----

// Walks the row streams of a banded layout as band_sweep_kernel does.
static void band_emulate(const BandPlan& P, const BandHost& H, const double* v,
                         std::vector<double>* out) {
    out->assign((size_t)P.S, 0.0);
    std::vector<double> acc((size_t)P.SB + 1);
    for (int sb = 0; sb < P.NSB; sb++)
        for (int part = 0; part < P.nparts; part++) {
            const int item = sb * P.nparts + part;
            const int vb0 = part * P.K;
            const int nk = std::min(P.NVB, vb0 + P.K) - vb0;
            const int rot = sb % nk;
            std::fill(acc.begin(), acc.end(), 0.0);
            for (int w = 0; w < P.NW; w++) {
                const int* rp = H.row_ptr.data() + ((size_t)item * P.NW + w) * (P.K + 1);
                for (int l = 0; l < 32; l++) {
                    double sum = 0.0;
                    for (int k = 0; k < nk; k++) {
                        int r = k + rot;
                        if (r >= nk) r -= nk;
                        const int vbase = (vb0 + r) * P.VB;
                        for (int row = rp[k]; row < rp[k + 1]; row++) {
                            const uint32_t* R = H.stream.data() + (size_t)row * 96;
                            const uint32_t key = R[l];
                            const double a = reinterpret_cast<const double*>(R + 32)[l];
                            const double vv = (key >> 16) == (uint32_t)P.SB ? 0.0
                                                                            : v[vbase + (key & 0x7fffu)];
                            sum += vv * a;
                            if (key & kBandLast) {
                                acc[key >> 16] += sum;
                                sum = 0.0;
                            }
                        }
                    }
                }
            }
            const int nseg = std::min(P.SB, P.S - sb * P.SB);
            for (int q = 0; q < nseg; q++) (*out)[(size_t)sb * P.SB + q] += acc[q];
        }
}

int ipxgpu_band_selftest(int64_t m, int64_t n, const int64_t* AIp, const int64_t* AIi,
                         const double* AIx, const double* x, int32_t force, double out[6]) {
    if (m <= 0 || n <= 0 || !AIp || !AIi || !AIx || !x || !out)
        return fail(IPXGPU_ERR_ARGUMENT, "invalid arguments");
    const int64_t nnz = AIp[n];
    if (nnz >= INT32_MAX || n >= INT32_MAX) return fail(IPXGPU_ERR_UNSUPPORTED, "too large");
    try {
        std::vector<int> cp((size_t)n + 1), ci((size_t)nnz);
        for (int64_t j = 0; j <= n; j++) cp[j] = (int)AIp[j];
        for (int64_t p = 0; p < nnz; p++) ci[p] = (int)AIi[p];
        std::vector<int> rp((size_t)m + 1, 0), rj((size_t)nnz);
        std::vector<double> rx((size_t)nnz);
        for (int64_t p = 0; p < nnz; p++) rp[ci[p] + 1]++;
        for (int64_t i = 0; i < m; i++) rp[i + 1] += rp[i];
        {
            std::vector<int> next(rp.begin(), rp.end() - 1);
            for (int j = 0; j < (int)n; j++)
                for (int p = cp[j]; p < cp[j + 1]; p++) {
                    const int put = next[ci[p]]++;
                    rj[put] = j;
                    rx[put] = AIx[p];
                }
        }
        std::vector<double> t_ref((size_t)n, 0.0), y_ref((size_t)m, 0.0);
        for (int j = 0; j < (int)n; j++)
            for (int p = cp[j]; p < cp[j + 1]; p++) t_ref[j] += x[ci[p]] * AIx[p];
        for (int j = 0; j < (int)n; j++)
            for (int p = cp[j]; p < cp[j + 1]; p++) y_ref[ci[p]] += t_ref[j] * AIx[p];
        const double ratio = force ? 1e30 : 1.5;
        for (int sweep = 0; sweep < 2; sweep++) {
            double* o = out + 3 * sweep;
            o[0] = o[1] = o[2] = 0.0;
            BandPlan P;
            const bool ok = sweep == 0 ? plan_band(&P, (int)m, (int)n, nnz, 148, ratio)
                                       : plan_band(&P, (int)n, (int)m, nnz, 148, ratio);
            if (!ok) continue;
            BandHost H;
            std::vector<int> spilled;  // dense segments: left out of the streams (sum 0)
            const bool built =
                sweep == 0 ? band_build(P, cp.data(), ci.data(), AIx, &H, kBandMaxRun, &spilled)
                           : band_build(P, rp.data(), rj.data(), rx.data(), &H, kBandMaxRun, &spilled);
            if (!built) continue;
            o[0] = spilled.empty() ? 1.0 : 2.0;
            std::vector<double> got;
            band_emulate(P, H, sweep == 0 ? x : t_ref.data(), &got);
            std::vector<double> ref = sweep == 0 ? t_ref : y_ref;
            for (int sgm : spilled) ref[sgm] = 0.0;
            double e = 0.0;
            for (size_t i = 0; i < ref.size(); i++) e = std::max(e, std::fabs(got[i] - ref[i]));
            o[1] = e;
            o[2] = H.rows > 0 ? (double)H.pad_entries / (32.0 * (double)H.rows) : 0.0;
        }
    } catch (const std::bad_alloc&) {
        return fail(IPXGPU_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    return IPXGPU_OK;
}

// ---- measurement helper ----

int ipxgpu_time_normal_apply(ipxgpu_ctx* c, int reps, int flush_l2, double out_ms[3]) {
    IPXGPU_TRY(check_ctx(c));
    if (!c->prepared) return fail(IPXGPU_ERR_STATE, "normal matrix not prepared");
    if (reps < 1 || !out_ms) return fail(IPXGPU_ERR_ARGUMENT, "invalid arguments");
    if (c->nranks != 1) return fail(IPXGPU_ERR_UNSUPPORTED, "timing helper is single-shard");
    if (flush_l2 && !c->flush_buf) {
        c->flush_bytes = (size_t)256 << 20;
        IPXGPU_TRY(dev_alloc(&c->flush_buf, c->flush_bytes));
    }
    cudaEvent_t e0, e1, e2;
    IPXGPU_CUDA(cudaEventCreate(&e0));
    IPXGPU_CUDA(cudaEventCreate(&e1));
    IPXGPU_CUDA(cudaEventCreate(&e2));
    double tot[3] = {0, 0, 0};
    const int np = (int)c->panels.size();
    int rc = IPXGPU_OK;
    for (int r = 0; r < reps && rc == IPXGPU_OK; r++) {
        if (flush_l2) cudaMemsetAsync(c->flush_buf, r & 0xff, c->flush_bytes, c->stream);
        float t1 = 0.f, t2 = 0.f;
        if (c->band1 || c->band2) {
            // A banded sweep is a whole-matrix stage: time the two stages.
            cudaEventRecord(e0, c->stream);
            if (c->band1) {
                BandArgs a1{c->xin, c->Wc, nullptr, nullptr, c->t, kApplyPlain, kSlotNone};
                rc = launch_band(c, *c->band1, a1, kBandColScale, nullptr);
                if (rc == IPXGPU_OK && c->spill1.nseg > 0) {
                    OpColDotScaleSpill op{c->xin, c->Wc, c->t, c->spill1.map};
                    rc = launch_sweep(c, op, c->spill1.tiles, c->spill1.A, nullptr);
                }
            } else {
                for (int k = 0; k < np && rc == IPXGPU_OK; k++) {
                    OpColDotScale op1{c->xin, c->Wc, c->t};
                    rc = launch_sweep(c, op1, c->panels[k].col_tiles, c->csc, nullptr);
                }
            }
            cudaEventRecord(e1, c->stream);
            if (rc == IPXGPU_OK && c->band2) {
                BandArgs a2{c->t, nullptr, c->Ws, c->xin, c->ybuf, kApplyPlain, kSlotNone};
                rc = launch_band(c, *c->band2, a2, kBandRowFinal, nullptr);
                if (rc == IPXGPU_OK && c->spill2.nseg > 0) {
                    OpRowGatherSpill op{c->t, c->xin, c->ybuf, c->spill2.map, (int)c->m,
                                        kApplyPlain, kSlotNone};
                    rc = launch_sweep(c, op, c->spill2.tiles, c->spill2.A, nullptr);
                }
            } else {
                for (int k = 0; k < np && rc == IPXGPU_OK; k++) {
                    OpRowGather op2{c->t, c->xin, c->Ws, c->ybuf, (int)c->m, k == 0, k == np - 1,
                                    kApplyPlain, kSlotNone};
                    rc = launch_sweep(c, op2, c->panels[k].row_tiles, c->panels[k].csr, nullptr);
                }
            }
            cudaEventRecord(e2, c->stream);
            cudaEventSynchronize(e2);
            cudaEventElapsedTime(&t1, e0, e1);
            cudaEventElapsedTime(&t2, e1, e2);
        }
        for (int k = 0; k < np && rc == IPXGPU_OK && !(c->band1 || c->band2); k++) {
            const Panel& P = c->panels[k];
            cudaEventRecord(e0, c->stream);
            OpColDotScale op1{c->xin, c->Wc, c->t};
            rc = launch_sweep(c, op1, P.col_tiles, c->csc, nullptr);
            cudaEventRecord(e1, c->stream);
            OpRowGather op2{c->t, c->xin, c->Ws, c->ybuf, (int)c->m, k == 0, k == np - 1,
                            kApplyPlain, kSlotNone};
            if (rc == IPXGPU_OK) rc = launch_sweep(c, op2, P.row_tiles, P.csr, nullptr);
            cudaEventRecord(e2, c->stream);
            cudaEventSynchronize(e2);
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, e0, e1);
            cudaEventElapsedTime(&b, e1, e2);
            t1 += a;
            t2 += b;
        }
        tot[1] += t1;
        tot[2] += t2;
        tot[0] += t1 + t2;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    for (int k = 0; k < 3; k++) out_ms[k] = tot[k] / reps;
    return rc;
}

}  // extern "C"

#include "split_api.inc"
#include "maxvol_api.inc"
#include "madd_api.inc"
