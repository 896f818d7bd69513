# e2e of km_find_text for a few share patterns of the sub-batches (KM_SUB_WEIGHTS; measurement aid)
A="--steps 20 --warmup 5 --no-lookup --no-cpu-baseline --no-tier2"
run() { KM_SUB_WEIGHTS=$2 python bench.py $A --n-sub $1 2>>gpurun_out/subw.err | python -c "
import sys, json
b = json.loads(sys.stdin.read()); print('n_sub $1 weights $2 e2e %.4f ms  identical %s' % (b['e2e']['ms_per_step'], b['e2e']['device_ms']['text_identical_to_one_call']))"; }
for w in "$@"; do n=$(echo $w | tr ',' '\n' | wc -l); run $n $w; done
