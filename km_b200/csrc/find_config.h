// Launch geometry of the find_mutation kernels: shared by the kernels' translation units (launch bounds,
// shared-memory classes) and by the host code that sizes a batch (plan_api.cu).
#pragma once
#include "graph.h"

namespace km {

#ifndef KM_CTA
#define KM_CTA 128
#endif

#ifndef KM_WALK_WARPS
#define KM_WALK_WARPS 4
#endif
#ifndef KM_WALK_MINB
#define KM_WALK_MINB 9       // (measured: 7 -> 0.271, 8 -> 0.262, 9 -> 0.256, 10 -> 0.296 ms) CTAs per SM the shared-memory walk's registers are budgeted for
#endif
// (measured: 4 warps per CTA and room for ~100 registers -- no spills -- beat 8 warps at 64 registers, 0.307 vs 0.333 ms)
#ifndef KM_PROBE_WARPS
#define KM_PROBE_WARPS 4
#endif
#ifndef KM_PROBE_MINB
#define KM_PROBE_MINB 5
#endif
#ifndef KM_PROBE_MINB_LINKED
#define KM_PROBE_MINB_LINKED 10
#endif
#ifndef KM_SMALL_NODES
#define KM_SMALL_NODES 512      // largest shared-memory class (graph nodes incl. the two caps)
#endif
// resident CTAs per SM the register allocation aims at: the 512-node class is held to 5 by its shared memory
#ifndef KM_GRAPH_SMALL_MINB
#define KM_GRAPH_SMALL_MINB 5
#endif
#ifndef KM_GRAPH_TINY_MINB
#define KM_GRAPH_TINY_MINB 8
#endif
// persistent CTAs per SM launched for each class (they take targets from a shared cursor)
#ifndef KM_GRAPH_SMALL_GRID
#define KM_GRAPH_SMALL_GRID 5
#endif
#ifndef KM_GRAPH_TINY_GRID
#define KM_GRAPH_TINY_GRID 10
#endif
#ifndef KM_TINY_NODES
#define KM_TINY_NODES 256
#endif
// the bubble pass (graph_bubble.h): warps per CTA, and the CTAs per SM its registers are budgeted for / that are launched
// (a warp's scratch is 8 KB in the 256-node class, 16 KB in the 512-node class)
#ifndef KM_BUBBLE_THREADS
#define KM_BUBBLE_THREADS 64     // threads per target: 32 = a warp (KM_BUBBLE_WARPS targets per CTA), 64 / 128 = a CTA
#endif
#ifndef KM_BUBBLE_WARPS
#define KM_BUBBLE_WARPS 4
#endif
#define KM_BUBBLE_SLOTS (KM_BUBBLE_THREADS == 32 ? KM_BUBBLE_WARPS : 1)      // targets (scratch areas) per CTA
#ifndef KM_BUBBLE_TINY_MINB
#define KM_BUBBLE_TINY_MINB 16
#endif
#ifndef KM_BUBBLE_SMALL_MINB
#define KM_BUBBLE_SMALL_MINB 10
#endif
#define KM_SMALL_CAND 64
#define KM_SMALL_PATHS 64
#define KM_SMALL_COLS 8


KM_HOSTDEV ScratchLayout class_layout(int nodes) {
    return make_layout(nodes - 2, KM_SMALL_CAND, KM_SMALL_PATHS, KM_SMALL_COLS, 1);
}

}  // namespace km
