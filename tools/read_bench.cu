// Pure-read bandwidth of B200 HBM for a few access shapes (tuning reference for
// the banded sweeps: what can a streaming read reach at all?).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2);} } while (0)

// grid-stride 16-byte loads, U independent loads per thread per trip
template <int U>
__global__ void __launch_bounds__(1024) read16(const uint4* p, size_t n16, unsigned* sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n16; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// per-warp private contiguous streams (like the banded sweep): warp w reads
// rows [w*R, (w+1)*R) of 384 bytes, 4 B + 8 B per lane per row, D rows in flight
template <int D>
__global__ void __launch_bounds__(1024) read_rows(const unsigned char* p, int rows_per_warp, unsigned* sink) {
    const int lane = threadIdx.x & 31;
    const size_t gw = (size_t)blockIdx.x * 32 + (threadIdx.x >> 5);
    const unsigned char* base = p + gw * rows_per_warp * 384 + lane * 4;
    double acc = 0;
    for (int r = 0; r < rows_per_warp; r += D) {
        unsigned k[D]; double a[D];
#pragma unroll
        for (int u = 0; u < D; u++) {
            k[u] = __ldcs((const unsigned*)(base + (size_t)(r + u) * 384));
            a[u] = __ldcs((const double*)(base + (size_t)(r + u) * 384 + 128 + lane * 4));
        }
#pragma unroll
        for (int u = 0; u < D; u++) acc += a[u] + __uint_as_float(k[u]);
    }
    if (acc == 1.2345e300) *sink = 1;
}

int main() {
    const size_t bytes = 1056ull << 20;
    unsigned char* buf; unsigned* sink; char* flush;
    CK(cudaMalloc(&buf, bytes + (1 << 20))); CK(cudaMemset(buf, 1, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&flush, 512ull << 20));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char* name, auto&& fn) {
        float best = 1e9;
        for (int it = 0; it < 8; it++) {
            CK(cudaMemsetAsync(flush, it, 512ull << 20));
            CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-44s %.1f us  %.0f GB/s\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9);
    };
    const size_t n16 = bytes / 16;
    timeit("read16 U=4 grid 148x1024", [&] { read16<4><<<148, 1024>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=8 grid 148x1024", [&] { read16<8><<<148, 1024>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=4 grid 296x1024", [&] { read16<4><<<296, 1024>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=4 grid 1184x256", [&] { read16<4><<<1184, 256>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=8 grid 2368x256", [&] { read16<8><<<2368, 256>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=8 grid 1184x256", [&] { read16<8><<<1184, 256>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=8 grid 592x256", [&] { read16<8><<<592, 256>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=16 grid 148x1024", [&] { read16<16><<<148, 1024>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=8 grid 148x512", [&] { read16<8><<<148, 512>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=2 grid 148x1024", [&] { read16<2><<<148, 1024>>>((const uint4*)buf, n16, sink); });
    timeit("read16 U=1 grid 148x1024", [&] { read16<1><<<148, 1024>>>((const uint4*)buf, n16, sink); });
    const int rpw = (int)(bytes / 384 / (148 * 32));
    timeit("rows D=4 148x1024 (warp streams)", [&] { read_rows<4><<<148, 1024>>>(buf, rpw, sink); });
    timeit("rows D=8 148x1024 (warp streams)", [&] { read_rows<8><<<148, 1024>>>(buf, rpw, sink); });
    timeit("rows D=2 148x1024 (warp streams)", [&] { read_rows<2><<<148, 1024>>>(buf, rpw, sink); });
    return 0;
}
