#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2z; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_gpu.time; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in 8 16; do
  RVB_REDUCE_WARPS=$w timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-sub-blocks > $O/C3_w$w.json 2> $O/C3_w$w.err
  python -c "
import json; d=json.loads(open('$O/C3_w$w.json').read().strip().splitlines()[-1]); print('C3 reduce warps $w ms/step', round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['stages_ms'].items()})"
done
