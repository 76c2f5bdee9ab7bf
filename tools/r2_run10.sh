#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
cd "$(dirname "$0")/.."
W=C2_shoebox_12tri_1Mx128_omni
timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > $O/c2.json 2> $O/c2.err
python -c "
import json; d=json.load(open('$O/c2.json')); print('C2 ms/step %.2f steady %.2f frac %.4f' % (d['ms_per_step'], d['stages_ms']['steady_phase'], d['roofline']['frac']), d['stages_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_bruteforce -s 7 -c 1 -o $O/prof_c2 python bench.py --workload $W --rays 262144 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_c2.log 2>&1
tail -2 $O/ncu_c2.log | head -c 300
