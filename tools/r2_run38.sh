#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2v; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bvh.py -q -x -k "visibility_map or pillar or dedupe or work_units" > $O/pytest_sel.log 2>&1; tail -5 $O/pytest_sel.log
timeout 300 python tools/wf_units_sweep.py 131072,262144,524288,1048576 0 2>&1 | grep rays | tee $O/sweep_auto.log
