// How much of an event-timed kernel is not the kernel's own threads?  Measurement aid (not part of the library).
// Kernels spin for a fixed time on %globaltimer; the event time minus that is launch + drain overhead for the
// launch shape march_kernel uses (147 CTAs x 448 threads, ~50 KB dynamic shared memory, totals to pinned memory).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void spin(unsigned long long ns, unsigned long long* host_out, unsigned long long* stamps)
{
    extern __shared__ unsigned s[];
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (threadIdx.x == 0 && stamps) stamps[blockIdx.x * 2] = t0;
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
    if (threadIdx.x == 0) s[0] = (unsigned)t;
    __syncthreads();
    if (threadIdx.x == 0 && stamps) stamps[blockIdx.x * 2 + 1] = t;
    if (host_out && blockIdx.x == 0 && threadIdx.x == 0) { *host_out = t; __threadfence_system(); }
}
__global__ void fill(unsigned char* p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1; }

static float med(std::vector<float> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

int main()
{
    unsigned char* flush; size_t fb = 256u << 20; cudaMalloc(&flush, fb);
    unsigned long long* host; cudaHostAlloc(&host, 64, cudaHostAllocMapped);
    unsigned long long* stamps; cudaMalloc(&stamps, 4096 * 16);
    cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int grid, block, smem; unsigned long long ns; int host; int do_flush; const char* name; };
    Cfg cfgs[] = {
        {147, 448, 50 * 1024, 0, 0, 1, "147x448 50KB spin 0, flush"},
        {147, 448, 50 * 1024, 0, 0, 0, "147x448 50KB spin 0, no flush"},
        {147, 448, 50 * 1024, 40000, 0, 1, "147x448 50KB spin 40us, flush"},
        {147, 448, 50 * 1024, 40000, 1, 1, "147x448 50KB spin 40us + pinned write, flush"},
        {147, 448, 0, 40000, 0, 1, "147x448 0KB spin 40us, flush"},
        {512, 128, 0, 40000, 0, 1, "512x128 0KB spin 40us, flush"},
        {147, 448, 50 * 1024, 40000, 0, 0, "147x448 50KB spin 40us, no flush"},
    };
    for (auto& c : cfgs) {
        std::vector<float> ts, span;
        for (int it = 0; it < 25; ++it) {
            if (c.do_flush) fill<<<1184, 256>>>(flush, fb);
            cudaEventRecord(e0);
            spin<<<c.grid, c.block, c.smem>>>(c.ns, c.host ? host : nullptr, stamps);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            std::vector<unsigned long long> h(c.grid * 2);
            cudaMemcpy(h.data(), stamps, c.grid * 16, cudaMemcpyDeviceToHost);
            unsigned long long lo = ~0ull, hi = 0;
            for (int b = 0; b < c.grid; ++b) { lo = std::min(lo, h[2 * b]); hi = std::max(hi, h[2 * b + 1]); }
            if (it >= 5) { ts.push_back(ms * 1e3f); span.push_back((hi - lo) * 1e-3f); }
        }
        printf("%-50s event %.2f us  first-start..last-end %.2f us  overhead %.2f us\n", c.name, med(ts), med(span), med(ts) - med(span));
    }
    return 0;
}
