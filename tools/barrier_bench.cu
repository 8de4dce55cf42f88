// Microbenchmark of grid-barrier variants for the persistent CR kernel
// (148 CTAs x 1024 threads, cooperative launch).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2);} } while (0)

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int V>
__device__ __forceinline__ void barrier(unsigned* count, unsigned gen) {
    if (V == 0 || V == 4) fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (V == 3) fence_proxy_async();
        if (V != 5) __threadfence();
        if (V == 5) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(count) : "memory");
        else asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(count) : "memory");
        const unsigned target = gen * gridDim.x;
        unsigned now;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(count) : "memory");
        } while ((int)(now - target) < 0);
        if (V != 2 && V != 5) __threadfence();
        if (V == 5) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        if (V == 3) fence_proxy_async();
    }
    __syncthreads();
    if (V == 0) fence_proxy_async();
}

template <int V>
__global__ void __launch_bounds__(1024, 1) bar_kernel(unsigned* count, int n, double* sink, unsigned long long* out) {
    unsigned long long t0 = 0;
    double acc = 0;
    for (int k = 1; k <= n; k++) {
        if (k == 11 && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        // a little work with global stores, as the CR phases have
        sink[(size_t)blockIdx.x * 1024 + threadIdx.x] = acc + k;
        barrier<V>(count, (unsigned)k);
        acc += sink[(size_t)((blockIdx.x + 1) % gridDim.x) * 1024 + threadIdx.x];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out[0] = t1 - t0;
    }
    if (acc == 12345.678) sink[0] = acc;
}

template <int V>
static void run(const char* name, int grid) {
    unsigned* count; double* sink; unsigned long long* out;
    CK(cudaMalloc(&count, 4)); CK(cudaMemset(count, 0, 4));
    CK(cudaMalloc(&sink, (size_t)grid * 1024 * 8)); CK(cudaMemset(sink, 0, (size_t)grid * 1024 * 8));
    CK(cudaMalloc(&out, 8));
    int n = 210;
    CK(cudaFuncSetAttribute(bar_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
    void* args[] = {&count, &n, &sink, &out};
    CK(cudaLaunchCooperativeKernel((void*)bar_kernel<V>, dim3(grid), dim3(1024), args, 180 * 1024, 0));
    CK(cudaDeviceSynchronize());
    unsigned long long ns;
    CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
    printf("%-52s %.2f us per barrier+work\n", name, ns / 200.0 / 1e3);
    cudaFree(count); cudaFree(sink); cudaFree(out);
}

int main() {
    int grid = 148;
    run<0>("0: proxy fences all threads, 2 threadfences", grid);
    run<1>("1: no proxy fences, 2 threadfences", grid);
    run<2>("2: no proxy fences, release fence only", grid);
    run<3>("3: proxy fences thread 0 only, 2 threadfences", grid);
    run<4>("4: proxy fence all threads before only", grid);
    run<5>("5: red.release + fence.acq_rel", grid);
    return 0;
}
