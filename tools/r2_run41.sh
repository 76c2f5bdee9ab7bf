#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2y; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_gpu.time; tail -3 $O/pytest_gpu.log
for v in 0 1; do
  RVB_BF_IMG_FULL=$v timeout 600 python bench.py --workload C2_shoebox_12tri_1Mx128_omni --steps 3 --no-cpu-baseline > $O/C2_full$v.json 2> $O/C2_full$v.err
  python -c "
import json; d=json.loads(open('$O/C2_full$v.json').read().strip().splitlines()[-1]); print('C2 full=$v ms/step', round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['stages_ms'].items()})"
done
timeout 600 python bench.py --workload C5_hall_64pairs_16kx128 --steps 3 --no-cpu-baseline > $O/C5.json 2> $O/C5.err
python -c "
import json; d=json.loads(open('$O/C5.json').read().strip().splitlines()[-1]); print('C5 ms/step', d['ms_per_step'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
