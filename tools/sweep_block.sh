#!/bin/bash
# sweep the brute-force CTA size (RVB_BLOCK) on the bench default and on the kernel-proper workload
for wl in C2_shoebox_12tri_1Mx128_omni X_room_4ktri_303kx64_bruteforce; do
  for b in 256 384 512 768 1024; do
    RVB_BLOCK=$b python bench.py --no-cpu-baseline --workload $wl --steps 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl', $b, 'ms/step %.2f' % d['ms_per_step'], 'steady %.2f ms' % d['stages_ms']['steady_phase'], 'image %.2f ms' % d['stages_ms']['image_phase'], 'frac %.3f' % d['roofline']['frac'])
"
  done
done
