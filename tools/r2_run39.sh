#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2mm; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bvh.py -q -x > $O/pytest_sel.log 2>&1; tail -3 $O/pytest_sel.log
RVB_TIMING=1 timeout 600 python bench.py --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/timing.json 2> $O/timing.log
tail -14 $O/timing.log | grep -E "micmap|image_phase|dedupe|trace.total"
timeout 600 python bench.py --workload C5_hall_64pairs_16kx128 --steps 3 --no-cpu-baseline > $O/C5.json 2> $O/C5.err
python -c "
import json; d=json.loads(open('$O/C5.json').read().strip().splitlines()[-1]); print('C5 ms/step', d['ms_per_step'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
