// Which lanes of a warp-wide 64-bit shared-memory access can conflict on B200? Times LDS.64 and
// LDS.64+STS.64 streams for a few lane -> address patterns (one CTA of 1024 threads per SM,
// like the banded sweep). The bank-aware dealing of band_build (band_sweep.cuh) assumes that a
// 64-bit access is served half-warp by half-warp (lanes 0-15, 16-31), 16 bank pairs each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bench tools/lds_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s\n", cudaGetErrorString(e_)); exit(2);} } while (0)

__global__ void __launch_bounds__(1024, 1) k(const int* pat, int npat, int iters, int rmw, long long* cycles, double* sink) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc = 0;
    for (int p = 0; p < npat; p++) {
        const int base = pat[p * 32 + lane];
        __syncthreads();
        const long long t0 = clock64();
        int off = 0;
#pragma unroll 8
        for (int it = 0; it < iters; it++) {
            const int a = (base + off) & 16383;
            const double v = sm[a];
            if (rmw) sm[a] = v + 1.0; else acc += v;
            off += 16;  // keeps the bank pattern, moves the line
        }
        __syncthreads();
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) cycles[p] = t1 - t0;
    }
    if (acc == 1.2345) sink[0] = acc;
}

int main() {
    const char* names[] = {"lane (contiguous)", "lane % 16 (halves alike)", "(lane % 16) * 16 (one pair)",
                           "lane * 2 (2 per pair in a half)", "lane % 8 (quarters alike)",
                           "(lane%8) + 8*(lane>=16) (q0=q1, q2=q3 shifted)", "random", "random, distinct in each half",
                           "lanes 0-7,16-23 distinct; 8-15,24-31 repeat", "lane % 4"};
    const int NP = 10;
    int h[NP * 32];
    unsigned s = 12345;
    for (int l = 0; l < 32; l++) {
        h[0 * 32 + l] = l;
        h[1 * 32 + l] = l % 16;
        h[2 * 32 + l] = (l % 16) * 16;
        h[3 * 32 + l] = l * 2;
        h[4 * 32 + l] = l % 8;
        h[5 * 32 + l] = (l % 8) + 8 * (l >= 16);
        s = s * 1664525u + 1013904223u;
        h[6 * 32 + l] = (s >> 8) % 8192;
        h[9 * 32 + l] = l % 4;
    }
    for (int half = 0; half < 2; half++) {  // random permutation of the 16 pairs + random line
        int perm[16];
        for (int i = 0; i < 16; i++) perm[i] = i;
        for (int i = 15; i > 0; i--) { s = s * 1664525u + 1013904223u; int j = (s >> 8) % (i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
        for (int i = 0; i < 16; i++) { s = s * 1664525u + 1013904223u; h[7 * 32 + half * 16 + i] = perm[i] + 16 * ((s >> 8) % 500); }
    }
    for (int l = 0; l < 32; l++) h[8 * 32 + l] = (l & 7) + 8 * ((l >> 4) & 1);
    int* d_pat; long long* d_cyc; double* d_sink;
    CK(cudaMalloc(&d_pat, sizeof h)); CK(cudaMemcpy(d_pat, h, sizeof h, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_cyc, NP * 8)); CK(cudaMalloc(&d_sink, 8));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    const int iters = 2048;
    for (int rmw = 0; rmw < 2; rmw++) {
        for (int w = 0; w < 2; w++) { k<<<148, 1024, 16384 * 8>>>(d_pat, NP, iters, rmw, d_cyc, d_sink); CK(cudaDeviceSynchronize()); }
        long long c[NP];
        CK(cudaMemcpy(c, d_cyc, sizeof c, cudaMemcpyDeviceToHost));
        printf("%s: cycles per warp-wide access per SM (32 warps; 1 wavefront = 1 cycle)\n", rmw ? "LDS.64 + STS.64" : "LDS.64");
        for (int p = 0; p < NP; p++) printf("  %-52s %.2f\n", names[p], (double)c[p] / iters / 32.0);
    }
    return 0;
}
