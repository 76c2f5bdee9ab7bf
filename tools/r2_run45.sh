#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_gpu.time; tail -3 $O/pytest_gpu.log
timeout 300 python tools/wf_units_sweep.py 131072,262144,1048576 0 2>&1 | grep rays | tee $O/sweep.log
