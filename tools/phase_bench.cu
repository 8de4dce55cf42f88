// How fast can ONE phase of a persistent kernel stream ~120 MB? 148 CTAs x 1024
// threads alternate between two regions (so nothing stays in L2) with a grid
// barrier after every phase; CTA 0 stamps the phase times.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2);} } while (0)

__device__ __forceinline__ void barrier(unsigned* count, unsigned gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(count) : "memory");
        const unsigned target = gen * gridDim.x;
        unsigned now;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(count) : "memory");
        } while ((int)(now - target) < 0);
        __threadfence();
    }
    __syncthreads();
}

template <int D, int SMEM_KB>
__global__ void __launch_bounds__(1024, 1) phases(const unsigned char* buf, size_t region_bytes, int rows_per_warp,
                                                  int nphase, unsigned* count, unsigned long long* out, double* sink) {
    extern __shared__ unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const size_t gw = (size_t)blockIdx.x * 32 + (threadIdx.x >> 5);
    double acc = 0;
    for (int ph = 0; ph < nphase; ph++) {
        unsigned long long t0 = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const unsigned char* base = buf + (size_t)(ph & 1) * region_bytes + gw * rows_per_warp * 384 + lane * 4;
        unsigned ka[D], kb[D]; double aa[D], ab[D];
#pragma unroll
        for (int u = 0; u < D; u++) {
            ka[u] = __ldcs((const unsigned*)(base + (size_t)u * 384));
            aa[u] = __ldcs((const double*)(base + (size_t)u * 384 + 128 + lane * 4));
        }
        for (int r = 0; r < rows_per_warp; r += 2 * D) {
#pragma unroll
            for (int u = 0; u < D; u++) {
                kb[u] = __ldcs((const unsigned*)(base + (size_t)(r + D + u) * 384));
                ab[u] = __ldcs((const double*)(base + (size_t)(r + D + u) * 384 + 128 + lane * 4));
            }
#pragma unroll
            for (int u = 0; u < D; u++) acc += aa[u] + __uint_as_float(ka[u]);
#pragma unroll
            for (int u = 0; u < D; u++) {
                ka[u] = __ldcs((const unsigned*)(base + (size_t)(r + 2 * D + u) * 384));
                aa[u] = __ldcs((const double*)(base + (size_t)(r + 2 * D + u) * 384 + 128 + lane * 4));
            }
#pragma unroll
            for (int u = 0; u < D; u++) acc += ab[u] + __uint_as_float(kb[u]);
        }
        unsigned long long t1 = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        barrier(count, (unsigned)(ph + 1));
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t2;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
            out[2 * ph] = t1 - t0;
            out[2 * ph + 1] = t2 - t0;
        }
    }
    if (acc == 1.2345e300) sink[0] = acc + smem[0];
}

// The same with the consumer structure of the banded sweep: warp 31 of every CTA idles, the other
// warps stream rows [first[w], first[w+1]) of one contiguous region (ragged counts).
template <int D, int NB, int G = 0, int WORK = 0>
__global__ void __launch_bounds__(1024, 1) phases_ragged(const unsigned char* buf, size_t region_bytes,
                                                         const int* first, int nphase, unsigned* count,
                                                         unsigned long long* out, double* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0;
    for (int ph = 0; ph < nphase; ph++) {
        unsigned long long t0 = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        if (warp < 31) {
            const int gw = blockIdx.x * 31 + warp;
            const int r_begin = first[gw], r_end = first[gw + 1];
            const unsigned char* base = buf + (size_t)(ph & 1) * region_bytes + lane * 4;
            unsigned k[NB][D]; double a[NB][D];
#pragma unroll
            for (int b = 0; b < NB - 1; b++)
#pragma unroll
                for (int u = 0; u < D; u++) {
                    if (G && r_begin + b * D + u >= r_end) { k[b][u] = 0; a[b][u] = 0; continue; }
                    k[b][u] = __ldcs((const unsigned*)(base + (size_t)(r_begin + b * D + u) * 384));
                    a[b][u] = __ldcs((const double*)(base + (size_t)(r_begin + b * D + u) * 384 + 128 + lane * 4));
                }
            for (int r = r_begin; r < r_end; r += NB * D) {
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    const int nb = (b + NB - 1) % NB;  // buffer to refill: the one consumed last
#pragma unroll
                    for (int u = 0; u < D; u++) {
                        const size_t rr = (size_t)(r + (b + NB - 1) * D + u);
                        if (G && rr >= (size_t)r_end) { k[nb][u] = 0; a[nb][u] = 0; continue; }
                        k[nb][u] = __ldcs((const unsigned*)(base + rr * 384));
                        a[nb][u] = __ldcs((const double*)(base + rr * 384 + 128 + lane * 4));
                    }
#pragma unroll
                    for (int u = 0; u < D; u++)
                        if (r + b * D + u < r_end) {
                            double v = a[b][u] + __uint_as_float(k[b][u]);
                            if (WORK) {
                                // a dependent chain through shared memory, like gather + accumulate
                                extern __shared__ double sm[];
#pragma unroll
                                for (int q = 0; q < WORK; q++) {
                                    const int slot = (k[b][u] + q * 977 + lane * 131) & 8191;
                                    v += sm[slot];
                                    sm[(slot * 7 + 3) & 8191] = v;
                                }
                            }
                            acc += v;
                        }
                }
            }
        }
        unsigned long long t1 = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        barrier(count, (unsigned)(ph + 1));
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t2;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
            out[2 * ph] = t1 - t0;
            out[2 * ph + 1] = t2 - t0;
        }
    }
    if (acc == 1.2345e300) sink[0] = acc;
}

template <int D, int NB, int G = 0, int WORK = 0>
static void run_ragged(const unsigned char* buf, size_t region, int mean_rows, int spread, const char* name) {
    unsigned* count; unsigned long long* out; double* sink; int* d_first;
    const int nw = 148 * 31;
    int* first = (int*)malloc((nw + 1) * sizeof(int));
    unsigned s = 12345u;
    first[0] = 0;
    for (int w = 0; w < nw; w++) {
        s = s * 1664525u + 1013904223u;
        const int rows = mean_rows + (spread ? (int)((s >> 16) % (2 * spread + 1)) - spread : 0);
        first[w + 1] = first[w] + rows;
    }
    CK(cudaMalloc(&d_first, (nw + 1) * sizeof(int)));
    CK(cudaMemcpy(d_first, first, (nw + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&count, 4)); CK(cudaMemset(count, 0, 4));
    CK(cudaMalloc(&out, 64 * 8)); CK(cudaMalloc(&sink, 8));
    int nphase = 12;
    void* args[] = {&buf, &region, &d_first, &nphase, &count, &out, &sink};
    CK(cudaFuncSetAttribute(phases_ragged<D, NB, G, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 186 * 1024));
    CK(cudaLaunchCooperativeKernel((void*)phases_ragged<D, NB, G, WORK>, dim3(148), dim3(1024), args, 186 * 1024, 0));
    CK(cudaDeviceSynchronize());
    unsigned long long h[64];
    CK(cudaMemcpy(h, out, 24 * 8, cudaMemcpyDeviceToHost));
    double a = 0, b = 0;
    for (int ph = 4; ph < 12; ph++) { a += h[2 * ph]; b += h[2 * ph + 1]; }
    const double bytes = (double)first[nw] * 384;
    printf("%-44s stream %.1f MB: CTA0 own %.2f us, with barrier %.2f us -> %.0f GB/s\n", name, bytes / 1e6,
           a / 8e3, b / 8e3, bytes / (b / 8 * 1e-9) / 1e9);
    cudaFree(count); cudaFree(out); cudaFree(sink); cudaFree(d_first); free(first);
}

template <int D, int SMEM_KB>
static void run(const unsigned char* buf, size_t region, int rpw, const char* name) {
    unsigned* count; unsigned long long* out; double* sink;
    CK(cudaMalloc(&count, 4)); CK(cudaMemset(count, 0, 4));
    CK(cudaMalloc(&out, 64 * 8)); CK(cudaMalloc(&sink, 8));
    int nphase = 12;
    CK(cudaFuncSetAttribute(phases<D, SMEM_KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KB * 1024));
    void* args[] = {&buf, &region, &rpw, &nphase, &count, &out, &sink};
    CK(cudaLaunchCooperativeKernel((void*)phases<D, SMEM_KB>, dim3(148), dim3(1024), args, SMEM_KB * 1024, 0));
    CK(cudaDeviceSynchronize());
    unsigned long long h[64];
    CK(cudaMemcpy(h, out, 24 * 8, cudaMemcpyDeviceToHost));
    double a = 0, b = 0;
    for (int ph = 4; ph < 12; ph++) { a += h[2 * ph]; b += h[2 * ph + 1]; }
    const double bytes = (double)rpw * 384 * 148 * 32;
    printf("%-34s stream %.1f MB: CTA0 own %.2f us, with barrier %.2f us -> %.0f GB/s\n", name, bytes / 1e6,
           a / 8e3, b / 8e3, bytes / (b / 8 * 1e-9) / 1e9);
    cudaFree(count); cudaFree(out); cudaFree(sink);
}

int main() {
    const int rpw = 72;  // rows per warp and phase: 72 * 384 * 4736 = 131 MB
    const size_t region = (size_t)rpw * 384 * 148 * 32 + (1 << 20);
    unsigned char* buf;
    CK(cudaMalloc(&buf, 2 * region + (1 << 20)));
    CK(cudaMemset(buf, 1, 2 * region));
    run<4, 186>(buf, region, rpw, "D=4 smem 186 KB");
    run<4, 100>(buf, region, rpw, "D=4 smem 100 KB");
    run<4, 16>(buf, region, rpw, "D=4 smem 16 KB");
    run<8, 186>(buf, region, rpw, "D=8 smem 186 KB");
    run<2, 186>(buf, region, rpw, "D=2 smem 186 KB");
    run<12, 16>(buf, region, rpw, "D=12 smem 16 KB");
    // 31 consumer warps per CTA as in the banded sweep: 74 rows per warp = 131 MB
    run_ragged<4, 2>(buf, region, 74, 0, "31 warps, 74 rows each, D=4 x 2 batches");
    run_ragged<4, 2>(buf, region, 74, 4, "31 warps, 74 +- 4 rows, D=4 x 2 batches");
    run_ragged<4, 2>(buf, region, 74, 12, "31 warps, 74 +- 12 rows, D=4 x 2 batches");
    run_ragged<4, 3>(buf, region, 74, 4, "31 warps, 74 +- 4 rows, D=4 x 3 batches");
    run_ragged<4, 4>(buf, region, 74, 4, "31 warps, 74 +- 4 rows, D=4 x 4 batches");
    run_ragged<2, 4>(buf, region, 74, 4, "31 warps, 74 +- 4 rows, D=2 x 4 batches");
    run_ragged<8, 2>(buf, region, 74, 4, "31 warps, 74 +- 4 rows, D=8 x 2 batches");
    run_ragged<4, 2, 1>(buf, region, 74, 4, "guarded loads, D=4 x 2 batches");
    run_ragged<4, 3, 1>(buf, region, 74, 4, "guarded loads, D=4 x 3 batches");
    run_ragged<4, 4, 1>(buf, region, 74, 4, "guarded loads, D=4 x 4 batches");
    run_ragged<2, 4, 1>(buf, region, 74, 4, "guarded loads, D=2 x 4 batches");
    run_ragged<2, 6, 1>(buf, region, 74, 4, "guarded loads, D=2 x 6 batches");
    run_ragged<4, 2, 1, 1>(buf, region, 74, 4, "guarded, D=4 x 2, smem work 1/row");
    run_ragged<4, 2, 1, 2>(buf, region, 74, 4, "guarded, D=4 x 2, smem work 2/row");
    run_ragged<4, 3, 1, 2>(buf, region, 74, 4, "guarded, D=4 x 3, smem work 2/row");
    run_ragged<4, 4, 1, 2>(buf, region, 74, 4, "guarded, D=4 x 4, smem work 2/row");
    run_ragged<2, 6, 1, 2>(buf, region, 74, 4, "guarded, D=2 x 6, smem work 2/row");
    return 0;
}
