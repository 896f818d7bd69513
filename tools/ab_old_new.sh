A="--steps 2 --warmup 3 --no-lookup --no-cpu-baseline --no-tier2"
for v in old new; do
  if [ $v = old ]; then export KM_B200_LIB=tools/variants/libkm_b200_old.so; else unset KM_B200_LIB; fi
  KM_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"km_graph_bubble|km_graph_kernel|km_walk_small|km_ref_probe|km_schedule" --launch-skip 16 -c 24 --csv --log-file gpurun_out/r3u_launches_$v.csv python bench.py $A > gpurun_out/r3u_$v.log 2>&1
done
