// The device formatter's kernels (format.h) and their launcher: the text of `km find_mutation` built on the GPU.
#include <cuda_runtime.h>

#define KM_FORMAT_KERNELS 1
#include "format.h"
#include "find_launch.h"

using namespace km;

cudaError_t km_launch_format(const WalkView& W, const ResultView& R, const FormatView& F, int k, cudaStream_t s) {
    const int n = W.n_targets;
    km_format_measure_kernel<<<(n + 3) / 4, 128, 0, s>>>(W, R, F, k);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    km_format_scan_kernel<<<1, 1024, 0, s>>>(F, n);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    km_format_write_kernel<<<(n + 3) / 4, 128, 0, s>>>(W, R, F, k);
    return cudaGetLastError();
}
