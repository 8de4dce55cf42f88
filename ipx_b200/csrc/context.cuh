// Host-side context of libipxgpu: device-resident matrix layouts, work
// vectors and launch helpers.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "smw.cuh"
#include "../../include/ipxgpu.h"

namespace ipxgpu {

extern thread_local std::string g_last_error;

inline int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define IPXGPU_CUDA(call)                                                              \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? IPXGPU_ERR_OUT_OF_MEMORY   \
                                                          : IPXGPU_ERR_CUDA;           \
            return fail(code_, std::string(#call) + ": " + cudaGetErrorString(e_));    \
        }                                                                              \
    } while (0)

#define IPXGPU_TRY(call)              \
    do {                              \
        int rc_ = (call);             \
        if (rc_ != IPXGPU_OK) return rc_; \
    } while (0)

// Device copy of a tiling of one compressed structure.
struct TileSet {
    Tile* tiles = nullptr;
    int ntiles = 0;
    int num_long = 0;
    int* long_first = nullptr;
    double* long_partials = nullptr;
    unsigned* long_counters = nullptr;
    double* long_dots = nullptr;
    LongInfo info() const {
        return LongInfo{long_first, long_partials, long_counters, long_dots, num_long};
    }
};

// Compressed structure on the device (CSC or CSR), int32 indices.
struct DevMatrix {
    int nseg = 0;
    long long nnz = 0;
    int* ptr = nullptr;
    int* idx = nullptr;
    double* val = nullptr;
};

// Column panel: columns [c0, c1) of the shard. The panel's slice of the
// intermediate vector t (8*(c1-c0) bytes) is meant to stay in L2 between the
// column sweep that writes it and the row sweep that gathers from it.
struct Panel {
    int c0 = 0, c1 = 0;
    TileSet col_tiles;  // over columns [c0, c1) of the shard CSC
    DevMatrix csr;      // rows of A[:, c0:c1], column ids local to the shard
    TileSet row_tiles;
};


// Segments the banded layout of a sweep leaves out (dense rows / columns: runs longer than a
// lane can carry): a compact compressed copy swept by the generic kernel right after the banded
// one; map[k] = segment k's index in the full structure.
struct Spill {
    int nseg = 0;
    DevMatrix A;
    TileSet tiles;
    int* map = nullptr;
};

template <class T>
inline int dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    IPXGPU_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    return IPXGPU_OK;
}

template <class T>
inline void dev_free(T*& p) {
    if (p) cudaFree((void*)p);
    p = nullptr;
}

template <class T>
inline int upload(T** dst, const std::vector<T>& src, cudaStream_t s) {
    IPXGPU_TRY(dev_alloc(dst, src.size()));
    if (!src.empty())
        IPXGPU_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T),
                                    cudaMemcpyHostToDevice, s));
    return IPXGPU_OK;
}

struct SplitOperator;  // split.cuh
struct BandDev;        // band_sweep.cuh
struct Top2;           // maxvol.cuh

}  // namespace ipxgpu

struct ipxgpu_group;  // ipxgpu.cu: the contexts of a one-process, several-GPU group

struct ipxgpu_ctx {
    ipxgpu_group* group = nullptr;  // set in a group's handle (which holds no device data)
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t m = 0, n = 0;
    int rank = 0, nranks = 1;
    int64_t col_begin = 0, col_end = 0;
    int nloc = 0;  // structural columns held by this shard
    int64_t launches = 0;
    int num_sms = 148;

    ipxgpu::DevMatrix csc;  // shard's structural columns, rows global
    std::vector<ipxgpu::Panel> panels;

    // weights
    double* W_own = nullptr;     // nloc + m owned copy: [structural shard | slack]
    const double* Wc = nullptr;  // structural weights of the shard or nullptr (= 1)
    const double* Ws = nullptr;  // slack weights or nullptr (= 0)
    bool prepared = false;
    double* W_full = nullptr;    // n+m, kktdiag path (nranks == 1 only)
    double* resscale_kkt = nullptr;  // m
    bool kkt_factorized = false;

    double* t = nullptr;     // nloc
    double* xin = nullptr;   // m
    double* ybuf = nullptr;  // m+1
    double* diag = nullptr;  // m
    bool diag_ready = false;
    // dense-column part of the preconditioner (smw.cuh); inactive: pure diagonal
    ipxgpu::SmwDev smw;
    bool smw_active = false;
    double* W_mask = nullptr;  // nloc: structural weights with the dense columns zeroed

    // CR work vectors (allocated on first solve)
    double *v_y = nullptr, *v_r = nullptr, *v_s = nullptr, *v_p = nullptr, *v_Cp = nullptr,
           *v_Cs = nullptr, *v_q = nullptr, *v_rhs = nullptr, *v_resscale = nullptr,
           *v_hist = nullptr;
    int64_t hist_cap = 0;
    double* nvec[4] = {nullptr, nullptr, nullptr, nullptr};  // n+m scratch (kktdiag path)

    ipxgpu::Reduce red{nullptr, nullptr};
    int red_cap = 0;
    ipxgpu::CrState* st_dev = nullptr;
    ipxgpu::HostMirror* mirror_host = nullptr;
    ipxgpu::HostMirror* mirror_dev = nullptr;
    double* scalars = nullptr;  // small device scratch (8 doubles)

    // NCCL (loaded on demand)
    void* nccl_comm = nullptr;

    // basis path
    ipxgpu::SplitOperator* split = nullptr;
    ulonglong2* tri_ll = nullptr;   // m records {generation | x halves} of the sync-free solves
    unsigned tri_gen = 0;           // generation of the current solve
    unsigned* tri_err = nullptr;    // set by a solve that gave up waiting for a dependency
    int tri_grid = 0;
    unsigned long long* tri_trace = nullptr;  // tuning only (option "tri_trace")
    int tri_reference_order = 0;       // triangular solves: 1 = reference's summation order (option)

    // Maxvolume column sweeps (maxvol.cuh); held for the duration of a run
    double* mv_colscale = nullptr;    // n+m
    double* mv_colweights = nullptr;  // n+m
    double* mv_vec = nullptr;         // m: work / btran of the current sweep
    ipxgpu::Top2* mv_partials = nullptr;
    ipxgpu::Top2* mv_out = nullptr;
    unsigned* mv_ticket = nullptr;
    int mv_grid = 0;

    // banded shared-memory sweeps of the normal-matrix apply (may be null)
    ipxgpu::BandDev* band1 = nullptr;  // t = W .* (A'x): gather x, segments = columns
    ipxgpu::BandDev* band2 = nullptr;  // y = A t: gather t, segments = rows
    ipxgpu::Spill spill1, spill2;      // segments left to the generic kernel (usually none)

    // peer exchange (NVLink P2P) of the persistent CR kernel for sharded contexts
    void* xchg = nullptr;             // own exchange buffer: y[2][xchg_mpad] doubles, then flags
    size_t xchg_mpad = 0;
    size_t xchg_ll_off = 0;           // byte offset of the push-exchange records in xchg
    double* xchg_abort = nullptr;     // set by a stand-alone exchange that gave up waiting
    void* peer_base[16] = {nullptr};  // every rank's exchange buffer as mapped here (own: xchg)
    double** peer_dev = nullptr;      // device copy of peer_base
    bool peers_ready = false;
    bool peers_direct = false;        // peer_base holds other devices' pointers of THIS process
    unsigned xchg_gen = 0;            // cross-GPU synchronisations performed so far

    // persistent CR kernel (pcr_fused.cuh): grid barrier words and per-CTA partials
    unsigned* fused_bar = nullptr;
    unsigned* fused_tickets = nullptr;  // per row block of sweep 2
    unsigned* fused_flags = nullptr;    // readiness flags: [sweep-1 items | CTAs]
    double* fused_red = nullptr;
    int fused_grid = 0;

    // L2 flush buffer for measurement helpers
    char* flush_buf = nullptr;
    size_t flush_bytes = 0;
};

namespace ipxgpu {
inline int grid_for(const ipxgpu_ctx* c, long long n) {
    long long g = (n + kBlock - 1) / kBlock;
    const long long cap = (long long)c->num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
}  // namespace ipxgpu
