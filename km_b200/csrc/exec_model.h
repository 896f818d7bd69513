// Execution-model shim shared by every device header in this directory.
//
// Compiled by nvcc (the product, libkm_b200.so) the stage functions are __device__ ONLY:
// there is no host-callable copy of the hot path in the shipped library and no CPU
// fallback.  Compiled by g++ (tests/emu only, never by km_b200/) the same source becomes a
// single-lane sequential program -- tid 0 of 1, barriers are no-ops, atomics are plain
// loads/stores -- so the kernel LOGIC can be checked against the oracle on a machine
// without a GPU (tests/test_emu_*.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KM_HD __device__ __forceinline__
#define KM_HOSTDEV __host__ __device__ inline
#define KM_DEVICE_BUILD 1
#else
#define KM_HD static inline
#define KM_HOSTDEV static inline
#define KM_HOST_EMU 1
#include <cmath>
#include <cstring>
#endif

namespace km {

// Phase timer (measurement only): lane 0 of each CTA adds the SM cycles it spent between marks to
// a global table, read back by km_debug_phase_cycles.  Compiled in when KM_PHASE_TIMERS is defined.
#if KM_DEVICE_BUILD && defined(KM_PHASE_TIMERS)
static __device__ unsigned long long km_phase_cycles[64];      // (one copy per translation unit: graph_kernels.cu reads its own)
#define KM_DEBUG_TARGETS 65536
static __device__ unsigned int km_target_cycles[KM_DEBUG_TARGETS];      // graph-pass cycles of each target (last launch)
struct PhaseTimer {
    long long t0;
    __device__ __forceinline__ PhaseTimer() {
#ifdef __CUDA_ARCH__
        t0 = clock64();
#endif
    }
    __device__ __forceinline__ void mark(int phase) {
#ifdef __CUDA_ARCH__
        if (threadIdx.x == 0) { const long long t1 = clock64(); atomicAdd(&km_phase_cycles[phase], (unsigned long long)(t1 - t0)); t0 = t1; }
#endif
    }
    // the same for a warp-per-target kernel: lane 0 of every warp reports
    __device__ __forceinline__ void mark_warp(int phase) {
#ifdef __CUDA_ARCH__
        if ((threadIdx.x & 31) == 0) { const long long t1 = clock64(); atomicAdd(&km_phase_cycles[phase], (unsigned long long)(t1 - t0)); t0 = t1; }
#endif
    }
    // one warp of a CTA-per-target kernel reports (the caller passes its lane)
    __device__ __forceinline__ void mark_warp0(int phase, int lane) {
#ifdef __CUDA_ARCH__
        if (lane == 0) { const long long t1 = clock64(); atomicAdd(&km_phase_cycles[phase], (unsigned long long)(t1 - t0)); t0 = t1; }
#endif
    }
};
#else
struct PhaseTimer {
    KM_HD void mark(int) {}
    KM_HD void mark_warp(int) {}
    KM_HD void mark_warp0(int, int) {}
};
#endif

#if KM_DEVICE_BUILD
// One CTA works on one target.
// `rot` rotates the warp numbering: the serial stretches of a target run on "warp 0, lane 0", and
// with 4-warp CTAs warp w of every CTA lands on scheduler w of the SM -- rotating by the target
// number spreads those stretches over the four schedulers instead of queueing them all on one.
struct CtaCtx {
    int rot = 0;
    KM_HD int tid() const { return (int)((threadIdx.x + 32u * (unsigned)rot) & (blockDim.x - 1u)); }   // blockDim is a power of two
    KM_HD int nt() const { return blockDim.x; }
    KM_HD void sync() const { __syncthreads(); }
    KM_HD int sync_or(int p) const { return __syncthreads_or(p); }
    KM_HD int sync_and(int p) const { return __syncthreads_and(p); }
};
// One warp works on one target (the walk: most of its life a target has one or two live lanes, so
// a warp per target keeps 4x more targets in flight per SM than a CTA per target).
struct WarpCtx {
    KM_HD int tid() const { return threadIdx.x & 31; }
    KM_HD int nt() const { return 32; }
    KM_HD void sync() const { __syncwarp(); }
    KM_HD int sync_or(int p) const { return __any_sync(0xFFFFFFFFu, p); }
    KM_HD int sync_and(int p) const { return __all_sync(0xFFFFFFFFu, p); }
};
// Sum over the lanes of the calling warp (all 32 lanes must call); the total is valid on the lane
// for which warp_leader() is true.  Used to turn one atomic per lane into one per warp.
KM_HD unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    return v;
}
KM_HD bool warp_leader() { return (threadIdx.x & 31) == 0; }
KM_HD unsigned long long warp_bcast64(unsigned long long v) { return __shfl_sync(0xFFFFFFFFu, v, 0); }      // lane 0's value
KM_HD uint32_t warp_min32(uint32_t v) { return __reduce_min_sync(0xFFFFFFFFu, v); }
KM_HD uint32_t warp_add32(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }      // the sum, on every lane (one REDUX)
KM_HD uint32_t warp_or32(uint32_t v) { return __reduce_or_sync(0xFFFFFFFFu, v); }
// lanes holding the same 64-bit value (all 32 lanes must call)
KM_HD uint32_t warp_match64(uint64_t v) { return __match_any_sync(0xFFFFFFFFu, (unsigned long long)v); }
KM_HD int warp_shfl32(int v, int src_lane) { return __shfl_sync(0xFFFFFFFFu, v, src_lane); }
KM_HD int warp_shfl_down32(int v, int delta) { return __shfl_down_sync(0xFFFFFFFFu, v, delta); }
KM_HD int popc32(uint32_t x) { return __popc(x); }
KM_HD void fence_block() { __threadfence_block(); }
KM_HD uint32_t load_shared_volatile32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
KM_HD uint8_t load_shared_volatile8(const uint8_t* p) { return *reinterpret_cast<const volatile uint8_t*>(p); }
KM_HD uint32_t atomic_cas32(uint32_t* p, uint32_t cmp, uint32_t val) { return atomicCAS(p, cmp, val); }
KM_HD int warp_index(const CtaCtx& c) { return c.tid() >> 5; }
KM_HD int warp_count(const CtaCtx&) { return blockDim.x >> 5; }
KM_HD int warp_index(const WarpCtx&) { return 0; }
KM_HD int warp_count(const WarpCtx&) { return 1; }
KM_HD uint64_t atomic_cas64(uint64_t* p, uint64_t cmp, uint64_t val) {
    return atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)cmp, (unsigned long long)val);
}
KM_HD uint32_t atomic_add32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
KM_HD int32_t atomic_addi32(int32_t* p, int32_t v) { return atomicAdd(p, v); }
KM_HD uint32_t atomic_min32(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
KM_HD uint32_t atomic_or32(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
KM_HD int32_t atomic_mini32(int32_t* p, int32_t v) { return atomicMin(p, v); }
KM_HD unsigned long long atomic_add64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
// loads that must observe other threads' atomics: go to L2, never a stale L1 line
KM_HD uint64_t load_cg64(const uint64_t* p) { return __ldcg(reinterpret_cast<const unsigned long long*>(p)); }
KM_HD uint32_t load_cg32(const uint32_t* p) { return __ldcg(p); }
KM_HD uint8_t load_cg8(const uint8_t* p) { return __ldcg(p); }
KM_HD int clz64(uint64_t x) { return __clzll((long long)x); }
KM_HD int clz32(uint32_t x) { return __clz((int)x); }
KM_HD int ffs32(uint32_t x) { return __ffs((int)x); }
KM_HD int ffs64(uint64_t x) { return __ffsll((long long)x); }
KM_HD uint64_t atomic_or64(uint64_t* p, uint64_t v) { return atomicOr(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }
KM_HD uint64_t mulhi64(uint64_t a, uint64_t b) { return __umul64hi(a, b); }
KM_HD float add_f32(float a, float b) { return __fadd_rn(a, b); }
// system-scope atomics: the target may be a PEER GPU's memory (a cohort shard mapped over NVLink); they are
// carried out at the home GPU's L2, so all GPUs inserting into one shard see one another
KM_HD uint64_t atomic_cas64_sys(uint64_t* p, uint64_t cmp, uint64_t val) {
    return atomicCAS_system(reinterpret_cast<unsigned long long*>(p), (unsigned long long)cmp, (unsigned long long)val);
}
KM_HD void atomic_add32_sys(uint32_t* p, uint32_t v) { atomicAdd_system(p, v); }
KM_HD void atomic_or64_sys(uint64_t* p, uint64_t v) { atomicOr_system(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }
KM_HD void store32_sys(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
KM_HD uint64_t load64_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
#else
struct CtaCtx {
    int tid() const { return 0; }
    int nt() const { return 1; }
    void sync() const {}
    int sync_or(int p) const { return p; }
    int sync_and(int p) const { return p; }
};
typedef CtaCtx WarpCtx;
KM_HD unsigned long long warp_sum64(unsigned long long v) { return v; }
KM_HD bool warp_leader() { return true; }
KM_HD unsigned long long warp_bcast64(unsigned long long v) { return v; }
KM_HD uint32_t warp_min32(uint32_t v) { return v; }
KM_HD uint32_t warp_add32(uint32_t v) { return v; }
KM_HD uint32_t warp_or32(uint32_t v) { return v; }
KM_HD uint32_t warp_match64(uint64_t) { return 1u; }
KM_HD int warp_shfl32(int v, int) { return v; }
KM_HD int warp_shfl_down32(int v, int) { return v; }
KM_HD int popc32(uint32_t x) { return __builtin_popcount(x); }
KM_HD void fence_block() {}
KM_HD uint32_t load_shared_volatile32(const uint32_t* p) { return *p; }
KM_HD uint8_t load_shared_volatile8(const uint8_t* p) { return *p; }
KM_HD uint32_t atomic_cas32(uint32_t* p, uint32_t cmp, uint32_t val) { uint32_t o = *p; if (o == cmp) *p = val; return o; }
KM_HD int warp_index(const CtaCtx&) { return 0; }
KM_HD int warp_count(const CtaCtx&) { return 1; }
KM_HD uint64_t atomic_cas64(uint64_t* p, uint64_t cmp, uint64_t val) { uint64_t o = *p; if (o == cmp) *p = val; return o; }
KM_HD uint32_t atomic_add32(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
KM_HD int32_t atomic_addi32(int32_t* p, int32_t v) { int32_t o = *p; *p = o + v; return o; }
KM_HD uint32_t atomic_min32(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
KM_HD uint32_t atomic_or32(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
KM_HD int32_t atomic_mini32(int32_t* p, int32_t v) { int32_t o = *p; if (v < o) *p = v; return o; }
KM_HD unsigned long long atomic_add64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
KM_HD uint64_t load_cg64(const uint64_t* p) { return *p; }
KM_HD uint32_t load_cg32(const uint32_t* p) { return *p; }
KM_HD uint8_t load_cg8(const uint8_t* p) { return *p; }
KM_HD int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
KM_HD int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
KM_HD int ffs32(uint32_t x) { return __builtin_ffs((int)x); }
KM_HD int ffs64(uint64_t x) { return __builtin_ffsll((long long)x); }
KM_HD uint64_t atomic_or64(uint64_t* p, uint64_t v) { uint64_t o = *p; *p = o | v; return o; }
KM_HD uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
KM_HD float add_f32(float a, float b) { volatile float r = a + b; return r; }
KM_HD uint64_t atomic_cas64_sys(uint64_t* p, uint64_t cmp, uint64_t val) { uint64_t o = *p; if (o == cmp) *p = val; return o; }
KM_HD void atomic_add32_sys(uint32_t* p, uint32_t v) { *p += v; }
KM_HD void atomic_or64_sys(uint64_t* p, uint64_t v) { *p |= v; }
KM_HD void store32_sys(uint32_t* p, uint32_t v) { *p = v; }
KM_HD uint64_t load64_sys(const uint64_t* p) { return *p; }
#endif

}  // namespace km
