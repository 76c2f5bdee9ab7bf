#!/bin/bash
for cfg in "8 296 0" "8 296 1" "8 444 1" "8 222 1" "16 148 1"; do
  set -- $cfg
  RVB_REDUCE_DENSE=1 RVB_REDUCE_WARPS=$1 RVB_REDUCE_GRID=$2 RVB_REDUCE_ASC=$3 python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('warps $1 grid $2 asc $3', 'ms/step %.2f' % d['ms_per_step'], 'reduce %.3f ms' % d['stages_ms']['reduce'])
"
done
