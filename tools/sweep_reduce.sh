#!/bin/bash
# sweep the dense reduce kernel's shape: warps per segment and resident CTAs (window of segments in flight)
for cfg in "8 0" "8 296" "8 148" "16 296" "16 148" "16 74"; do
  set -- $cfg
  RVB_REDUCE_DENSE=1 RVB_REDUCE_WARPS=$1 RVB_REDUCE_GRID=$2 python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('warps $1 grid $2', 'ms/step %.2f' % d['ms_per_step'], 'reduce %.3f ms' % d['stages_ms']['reduce'])
"
done
