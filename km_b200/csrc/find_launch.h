// Host-callable launchers of the find_mutation kernels.  The kernels live in their own translation units
// (walk_kernels.cu, graph_kernels.cu, format_kernels.cu) so that each family compiles -- and is tuned -- on its own;
// the host code (plan_api.cu) only sees these declarations.
#pragma once
#include <cuda_runtime.h>

#include "graph.h"

namespace km { struct FormatView; }

cudaError_t km_find_kernels_init();          // opt-in shared memory of the graph kernels (once per device)
cudaError_t km_launch_encode(const km::WalkView& W, cudaStream_t s);
cudaError_t km_launch_ref_probe(const km::TableView& T, const km::WalkView& W, const km::FindParams& P, cudaStream_t s);
cudaError_t km_launch_walks(const km::TableView& T, const km::WalkView& W, const km::FindParams& P, cudaStream_t s);
cudaError_t km_launch_schedule(const km::WalkView& W, const km::ResultView& R, cudaStream_t s);
// cls 0 / 1: the 256- / 512-node shared-memory classes, 2: the general pass (scratch in HBM); `list`: the scheduler's work
// list the pass reads (0..2: the classes' own, 3 / 4: what the bubble pass of class 0 / 1 handed on)
cudaError_t km_launch_graph(int cls, int list, int grid, const km::TableView& T, const km::WalkView& W, const km::ScratchLayout& SL,
                            const km::ResultView& R, cudaStream_t s);
// the simple bubbles of class cls (0 / 1), one warp per target (graph_bubble.h)
cudaError_t km_launch_bubble(int cls, int grid, const km::TableView& T, const km::WalkView& W, const km::ResultView& R, cudaStream_t s);
bool km_bubble_pass_enabled();
cudaError_t km_launch_format(const km::WalkView& W, const km::ResultView& R, const km::FormatView& F, int k, cudaStream_t s);
