// Level-scheduled sparse triangular solves and the basis-preconditioned
// operator C = I + inv(B) N N' inv(B')  (reference src/sparse_matrix.cc:224-311,
// src/splitted_normal_matrix.cc:18-117).
//
// The reference solves L and U column-wise (scatter) and U', L' row-wise
// (gather), all strictly sequentially. Here all four solves are in gather
// form: every row i reads already-solved entries x[k], so the rows of one
// dependency level are independent. Rows are bucketed by level on the host;
// runs of narrow levels are executed by one CTA that steps through them with
// __syncthreads(), wide levels by a full grid. Each row accumulates its
// updates in exactly the order the reference applies them (rounded product,
// then add/subtract), so the solves are bit-identical to the CPU loops.
#pragma once

#include <algorithm>
#include <vector>

#include "context.cuh"
#include "spmv_kernels.cuh"

namespace ipxgpu {

constexpr int kTriBlock = 1024;      // CTA of the merged-level kernel
constexpr int kTriWideRows = 4096;   // levels at least this wide get a grid

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

// ---- host side ----

static int ensure_reduce(ipxgpu_ctx* c, int grid);
static int launch_normal_apply_w(ipxgpu_ctx* c, const double* Wc, const double* Ws,
                                 const double* x, double* y, int mode, int slot, CrState* st);

static void free_tri(TriSystem* T) {
    dev_free(T->d.ptr);
    dev_free(T->d.idx);
    dev_free(T->d.val);
    dev_free(T->d.diag);
    dev_free(T->d.order);
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
    IPXGPU_TRY(upload(&T->d.level_ptr, lptr, s));
    IPXGPU_CUDA(cudaStreamSynchronize(s));
    return IPXGPU_OK;
}

static int launch_tri(ipxgpu_ctx* c, const TriSystem& T, double* x, const CrState* st) {
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
