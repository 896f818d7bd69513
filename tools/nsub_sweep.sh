#!/bin/bash
# e2e of km_find_text against the number of sub-batches in flight (bench.py --n-sub), one GPU
for n in "$@"; do
  python bench.py --steps 20 --warmup 5 --no-lookup --no-cpu-baseline --no-tier2 --n-sub $n 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('n_sub $n', 'device %.3f ms' % b['ms_per_step'], 'e2e %.3f ms' % b['e2e']['ms_per_step'])
"
done
