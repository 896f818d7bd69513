KM_NVCC_EXTRA="-DKM_SMALL_NODES=448 -DKM_GRAPH_SMALL_MINB=6" python bench.py --no-cpu-baseline --no-lookup --steps 10 > gpurun_out/s5_h_448.json 2>/dev/null
KM_NVCC_EXTRA="-DKM_SMALL_NODES=448 -DKM_GRAPH_SMALL_MINB=6 -DKM_TINY_NODES=224" python bench.py --no-cpu-baseline --no-lookup --steps 10 > gpurun_out/s5_h_448_224.json 2>/dev/null
KM_NVCC_EXTRA="-DKM_TINY_NODES=288" python bench.py --no-cpu-baseline --no-lookup --steps 10 > gpurun_out/s5_h_512_288.json 2>/dev/null
