#!/bin/bash
# kernel times of the standard panel for each library variant (tools/build_variant.py): one bench.py line per variant
for lib in "$@"; do
  tag=$(basename "$lib" .so)
  KM_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-lookup --no-cpu-baseline --no-tier2 2> gpurun_out/var_$tag.err | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('$tag', 'step %.4f ms' % b['ms_per_step'], b['kernels']['km_ref_probe_kernel_ms'], b['kernels']['km_walk_kernels_ms'], b['kernels']['km_graph_kernels_ms'], 'e2e %.3f' % b['e2e']['ms_per_step'])
" || tail -5 gpurun_out/var_$tag.err
done
