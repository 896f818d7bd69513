// Random 32-byte sector gather from a PEER GPU's memory over NVLink (cohort mode's remote probes):
// which load flavour does the fabric like?   nvcc -O3 -arch=sm_100a peer_gather.cu -o peer_gather
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint64_t mix64(uint64_t z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
template <int MODE>
__global__ void gather(const uint4* __restrict__ buf, uint64_t n_sectors, uint64_t n_loads, uint64_t seed, uint32_t* sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n_loads; base += stride * 4) {
        uint64_t s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = __umul64hi(mix64(seed + (base + i * stride) * 0x9E3779B97F4A7C15ull), n_sectors);
        uint4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4* p = buf + 2 * s[i];
            if (MODE == 0) {
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[i].x), "=r"(a[i].y), "=r"(a[i].z), "=r"(a[i].w) : "l"(p));
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b[i].x), "=r"(b[i].y), "=r"(b[i].z), "=r"(b[i].w) : "l"(p + 1));
            } else if (MODE == 1) {
                asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[i].x), "=r"(a[i].y), "=r"(a[i].z), "=r"(a[i].w) : "l"(p));
                asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b[i].x), "=r"(b[i].y), "=r"(b[i].z), "=r"(b[i].w) : "l"(p + 1));
            } else if (MODE == 2) {
                asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[i].x), "=r"(a[i].y), "=r"(a[i].z), "=r"(a[i].w) : "l"(p));
                asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b[i].x), "=r"(b[i].y), "=r"(b[i].z), "=r"(b[i].w) : "l"(p + 1));
            } else if (MODE == 4) {   // one 32-byte load
                asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(*(uint64_t*)&a[i].x), "=l"(*(uint64_t*)&a[i].z), "=l"(*(uint64_t*)&b[i].x), "=l"(*(uint64_t*)&b[i].z) : "l"(p));
            } else if (MODE == 5) {   // one 32-byte load, no L1 allocation
                asm volatile("ld.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(*(uint64_t*)&a[i].x), "=l"(*(uint64_t*)&a[i].z), "=l"(*(uint64_t*)&b[i].x), "=l"(*(uint64_t*)&b[i].z) : "l"(p));
            } else if (MODE == 6) {   // one 128-byte line per QUAD: lane j of the quad loads sector j (one coalesced request)
                const uint64_t sq = __shfl_sync(0xFFFFFFFFu, s[i], threadIdx.x & 28);
                const uint4* pl = buf + 2 * ((sq & ~3ull) + (threadIdx.x & 3));
                asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(*(uint64_t*)&a[i].x), "=l"(*(uint64_t*)&a[i].z), "=l"(*(uint64_t*)&b[i].x), "=l"(*(uint64_t*)&b[i].z) : "l"(pl));
            } else {   // one 16-byte load only
                asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[i].x), "=r"(a[i].y), "=r"(a[i].z), "=r"(a[i].w) : "l"(p));
                b[i] = a[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc ^= a[i].x ^ a[i].y ^ b[i].z ^ b[i].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int MODE> float run(const uint4* buf, uint64_t n_sectors, uint64_t n_loads, uint32_t* sink, int blocks) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        gather<MODE><<<blocks, 256>>>(buf, n_sectors, n_loads, 77 + it, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
    }
    return best;
}
int main(int argc, char** argv) {
    if (argc > 1) {   // before anything else touches the device
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1]));
        size_t g0 = 0; cudaDeviceGetLimit(&g0, cudaLimitMaxL2FetchGranularity);
        printf("early set %s -> %s, now %zu\n", argv[1], cudaGetErrorString(e), g0);
    }
    int n = 0; CK(cudaGetDeviceCount(&n));
    const uint64_t n_loads = 1ull << 26;
    uint32_t* sink;
    CK(cudaSetDevice(0)); CK(cudaMalloc(&sink, 4));
    const char* names[7] = {"ld.global.nc.L1::no_allocate 2x16B", "ld.global 2x16B", "ld.global.cg 2x16B", "ld.global 1x16B", "ld.global.v4.u64 1x32B", "ld.global.L1::no_allocate.v4.u64 1x32B", "quad: 4 lanes x 32B = one 128B line (G lane-loads/s; lines/s = this / 4)"};
    if (n >= 2) CK(cudaDeviceEnablePeerAccess(1, 0));
    size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n", gran);
    if (argc > 1) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1])); cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity); printf("set %s -> %s, now %zu\n", argv[1], cudaGetErrorString(e), gran); }
    for (int where = 0; where < (n >= 2 ? 2 : 1); ++where)
        for (uint64_t gib : {64ull}) {
            uint4* buf;
            const uint64_t bytes = gib << 30, n_sectors = bytes / 32;
            CK(cudaSetDevice(where)); CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes)); CK(cudaDeviceSynchronize());
            CK(cudaSetDevice(0));
            const int blocks = 148 * 8;
            float t[7] = {run<0>(buf, n_sectors, n_loads, sink, blocks), run<1>(buf, n_sectors, n_loads, sink, blocks), run<2>(buf, n_sectors, n_loads, sink, blocks),
                          run<3>(buf, n_sectors, n_loads, sink, blocks), run<4>(buf, n_sectors, n_loads, sink, blocks), run<5>(buf, n_sectors, n_loads, sink, blocks), run<6>(buf, n_sectors, n_loads, sink, blocks)};
            for (int m = 0; m < 7; ++m) printf("%s %2llu GiB  %-40s %7.2f G loads/s\n", where ? "peer " : "local", (unsigned long long)gib, names[m], n_loads / t[m] / 1e6);
            CK(cudaSetDevice(where)); CK(cudaFree(buf));
        }
    return 0;
}
