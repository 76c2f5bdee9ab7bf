#!/bin/bash
# sweep the BVH steady kernel's occupancy target (RVB_BVH_MINB) on the C3 hall
for b in 0 6 7 8 10 12; do
  RVB_BVH_MINB=$b python bench.py --no-cpu-baseline --workload C3_hall_100ktri_1Mx128_stereo --rays ${1:-262144} --steps 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('minb', $b, 'ms/step %.2f' % d['ms_per_step'], 'steady %.2f ms' % d['stages_ms']['steady_phase'], 'image %.2f ms' % d['stages_ms']['image_phase'])
"
done
