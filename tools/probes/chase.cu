// Dependent-load latency of random 32-byte reads over a large span (what one level of a walk pays):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/chase tools/probes/chase.cu && tools/probes/chase [GiB]
// One thread per CTA follows a chain whose next address depends on the loaded value; `warps` chains per SM run
// side by side to show how the latency grows with the number of walks in flight.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__global__ void chase(const uint4* __restrict__ buf, uint64_t n_sectors, int hops, uint64_t* out, long long* cycles, int lanes) {
    if ((int)(threadIdx.x & 31) >= lanes) return;
    uint64_t x = mix(blockIdx.x * 1024ull + threadIdx.x + 1);
    const long long t0 = clock64();
    for (int i = 0; i < hops; ++i) {
        const uint64_t s = __umul64hi(mix(x + i), n_sectors);
        uint64_t a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(buf + 2 * s));
        x ^= a + b + c + d;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (x == 0x1234567) out[0] = x;
}
int main(int argc, char** argv) {
    const uint64_t gib = argc > 1 ? atoll(argv[1]) : 64;
    const uint64_t bytes = gib << 30, n_sectors = bytes / 32;
    uint4* buf; uint64_t* out; long long* cyc;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, 8 * 148 * 64));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int hops = 2000;
    struct { int ctas_per_sm, threads, lanes; } cfg[] = {{1, 32, 1}, {1, 32, 4}, {1, 32, 32}, {8, 32, 4}, {8, 128, 4}, {8, 128, 8}, {8, 128, 32}, {8, 256, 32}};
    for (auto& c : cfg) {
        const int blocks = 148 * c.ctas_per_sm;
        chase<<<blocks, c.threads>>>(buf, n_sectors, 64, out, cyc, c.lanes);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        chase<<<blocks, c.threads>>>(buf, n_sectors, hops, out, cyc, c.lanes);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[1]; CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        const double chains = (double)blocks * (c.threads / 32) * c.lanes;
        printf("span %llu GiB  %d CTAs/SM x %d threads, %d lanes/warp chasing: %.0f ns/hop (%.0f cycles), %.2f G loads/s\n", (unsigned long long)gib,
               c.ctas_per_sm, c.threads, c.lanes, 1e6 * ms / hops, (double)h[0] / hops, chains * hops / ms / 1e6);
    }
    return 0;
}
