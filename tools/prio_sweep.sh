# e2e of km_find_text for stream-priority modes x share patterns x panel seeds (measurement aid)
A="--steps 20 --warmup 5 --no-lookup --no-cpu-baseline --no-tier2"
for off in 0 4; do for pm in 1 0 2; do for w in 2,3,4,5 1,1,1,1 3,4,4,3; do
  KM_LANE_PRIORITIES=$pm KM_SUB_WEIGHTS=$w python bench.py $A --panel-offset $off 2>>gpurun_out/prio.err | python -c "
import sys, json
b = json.loads(sys.stdin.read()); print('seed+$off priorities $pm weights $w e2e %.4f ms' % b['e2e']['ms_per_step'])"
done; done; done
