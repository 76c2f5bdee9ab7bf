#!/bin/bash
# few-candidates de-duplication kernel; visibility-map resolution against the ray count
cd "$(dirname "$0")/.."
O=gpurun_out/r2d; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_gpu.time; tail -3 $O/pytest_gpu.log
for cpt in 16 8 4 2; do
  echo "cells per triangle $cpt"
  RVB_MICMAP_CPT=$cpt timeout 300 python tools/wf_units_sweep.py 131072,1048576 0 2>&1 | grep rays | tee -a $O/sweep_cpt.log
done
RVB_TIMING=1 timeout 600 python bench.py --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/timing.json 2> $O/timing.log
tail -14 $O/timing.log
