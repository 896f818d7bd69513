# ncu launch lists (gpu__time_duration only) of the bench command with an OLD build of the library and with the current one, on
# the same box: tools/variants/libkm_b200_old.so is built beforehand from an earlier commit, e.g.
#   git worktree add /tmp/old <commit> && (cd /tmp/old && python -m km_b200.build) && cp /tmp/old/km_b200/libkm_b200.so tools/variants/libkm_b200_old.so
# (measurement aid; results: profiles/r3u_launches_*_session.csv)
A="--steps 2 --warmup 3 --no-lookup --no-cpu-baseline --no-tier2"
for v in old new; do
  if [ $v = old ]; then export KM_B200_LIB=tools/variants/libkm_b200_old.so; else unset KM_B200_LIB; fi
  KM_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"km_graph_bubble|km_graph_kernel|km_walk_small|km_ref_probe|km_schedule" --launch-skip 16 -c 24 --csv --log-file gpurun_out/r3u_launches_$v.csv python bench.py $A > gpurun_out/r3u_$v.log 2>&1
done
