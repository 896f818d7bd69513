// Kernels of stage 1 of find_mutation (sm_100a): the packing of the uploaded targets, level 0 of every walk
// flat over all reference k-mers (K3a), the shared-memory walk and the general walk (K3b) -- MutationFinder.__init__
// / __extend, km/utils/MutationFinder.py:87-165.  One translation unit per kernel family: each compiles on its own.
#include <cuda_runtime.h>

#include "walk_small.h"
#include "find_config.h"
#include "find_launch.h"

namespace km {

// Once per upload, one warp per target: letters -> codes in place (A0 C1 G2 T3, anything else 255; the
// general walk reads these) and the 2-bit packed copy the probe and shared-memory walk kernels read
// (16 bases per word, first base in the top bits, >= 2 zero words after each target).
__global__ void km_encode_kernel(uint8_t* seq, const int64_t* seq_off, uint32_t* pack, const int64_t* pack_off, uint8_t* pre_bad,
                                 int n_targets) {
    const int t = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= n_targets) return;
    const int64_t s0 = seq_off[t], w0 = pack_off[t];
    const int len = (int)(seq_off[t + 1] - s0), nw = (int)(pack_off[t + 1] - w0);
    bool bad = false;
    for (int w = lane; w < nw; w += 32) {
        uint32_t word = 0;
        for (int j = 0; j < 16; ++j) {
            const int pos = 16 * w + j;
            uint32_t c = 0;
            if (pos < len) {
                const uint8_t ch = seq[s0 + pos];
                c = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 255u;
                seq[s0 + pos] = (uint8_t)c;
                if (c > 3u) { bad = true; c = 0u; }
            }
            word = (word << 2) | c;
        }
        pack[w0 + w] = word;
    }
    bad = __any_sync(0xFFFFFFFFu, bad);
    if (lane == 0) pre_bad[t] = bad ? 1 : 0;
}


// ---- K3: walk, one WARP per target ----------------------------------------------------------
// Two launches: the shared-memory walk takes every target of ordinary size (walk_small.h); the
// general walk, whose per-target state lives in HBM, takes the targets the first one deferred.
// K3a: level 0 of every walk, one warp per 32 reference k-mers, flat over the batch
// (two instances: with the neighbour masks a lane makes ONE table read and almost never a second, so that instance is
// compiled for twice the warps per SM; without them five reads are in flight per lane and registers buy more than warps)
template <bool LINKED>
__global__ void __launch_bounds__(32 * KM_PROBE_WARPS, LINKED ? KM_PROBE_MINB_LINKED : KM_PROBE_MINB) km_ref_probe_kernel(TableView T, WalkView W, FindParams P) {
    WarpCtx ctx;
    const int ch = (int)blockIdx.x * KM_PROBE_WARPS + (int)(threadIdx.x >> 5);
    if (ch >= W.n_chunks) return;
    ref_probe_chunk<WarpCtx, LINKED>(ctx, T, W, P, W.chunk_target[ch], W.chunk_start[ch]);
}

// Persistent warps: each takes the next target from a cursor until none is left.  (One CTA per four targets held its slot
// until the LONGEST of its four walks was over -- walks differ several-fold, 30 to 180 levels -- and the slots of the three
// finished warps sat idle meanwhile.)
__global__ void __launch_bounds__(32 * KM_WALK_WARPS, KM_WALK_MINB) km_walk_small_kernel(TableView T, WalkView W, FindParams P) {
    __shared__ WalkSmall M[KM_WALK_WARPS];
    WarpCtx ctx;
    for (;;) {
        int t = 0;
        if ((threadIdx.x & 31) == 0) t = (int)atomicAdd(W.walk_cursor, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= W.n_targets) return;
        const TargetGeom g = target_geom(W, t, T.k);
        if (!(walk_small_fits(g) && walk_small_target(ctx, T, W, P, t, M[threadIdx.x >> 5]))) {
            // too long for the shared-memory state, or more novel nodes than it holds: the general walk, state in HBM, by the
            // same warp (rare; a kernel of its own for these cost 7 us per batch whether or not there was one)
            __syncwarp();
            if ((threadIdx.x & 31) == 0) { W.status[t] = 0; W.lookups[t] = 0; W.n_kept[t] = 0; }     // (also drops the probe's limit flag)
            __syncwarp();
            walk_target(ctx, T, W, P, t);
        }
        __syncwarp();
    }
}

}  // namespace km

using namespace km;

cudaError_t km_launch_encode(const WalkView& W, cudaStream_t s) {
    const int n = W.n_targets;
    if (n) km_encode_kernel<<<(n + 7) / 8, 256, 0, s>>>(const_cast<uint8_t*>(W.codes), W.seq_off, const_cast<uint32_t*>(W.pack), W.pack_off,
                                                       const_cast<uint8_t*>(W.pre_bad), n);
    return cudaGetLastError();
}
cudaError_t km_launch_ref_probe(const TableView& T, const WalkView& W, const FindParams& P, cudaStream_t s) {
    const int grid = (W.n_chunks + KM_PROBE_WARPS - 1) / KM_PROBE_WARPS;
    if (W.n_chunks && T.linked) km_ref_probe_kernel<true><<<grid, 32 * KM_PROBE_WARPS, 0, s>>>(T, W, P);
    else if (W.n_chunks) km_ref_probe_kernel<false><<<grid, 32 * KM_PROBE_WARPS, 0, s>>>(T, W, P);
    return cudaGetLastError();
}
cudaError_t km_launch_walks(const TableView& T, const WalkView& W, const FindParams& P, cudaStream_t s) {
    const int n = W.n_targets;
    static int sm_count[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sm_count[dev]) {
        int v = 0;
        if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        sm_count[dev] = v > 0 ? v : 1;
    }
    const int want = (n + KM_WALK_WARPS - 1) / KM_WALK_WARPS, resident = sm_count[dev] * KM_WALK_MINB;
    km_walk_small_kernel<<<want < resident ? want : resident, 32 * KM_WALK_WARPS, 0, s>>>(T, W, P);
    return cudaGetLastError();
}

// per-phase SM cycles of the walk kernels (KM_PHASE_TIMERS builds only; the counters are per translation unit)
extern "C" int km_debug_walk_cycles(unsigned long long* out64, int reset) {
#ifdef KM_PHASE_TIMERS
    if (out64 && cudaMemcpyFromSymbol(out64, km_phase_cycles, 64 * sizeof(unsigned long long)) != cudaSuccess) return -3;
    if (reset) { unsigned long long z[64] = {0}; if (cudaMemcpyToSymbol(km_phase_cycles, z, sizeof(z)) != cudaSuccess) return -3; }
    return 0;
#else
    (void)out64; (void)reset;
    return -1;
#endif
}
