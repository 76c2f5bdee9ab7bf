#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
cd "$(dirname "$0")/.."
RVB_DEBUG_SIGNALS=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c5_worker" > $O/pytest_c5.log 2>&1
grep -n "rvb\] fatal" -A 30 $O/pytest_c5.log | head -60
tail -3 $O/pytest_c5.log
( time timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_fullsize.py::test_c5_worker_threads_equal_sequential > $O/pytest_rest.log 2>&1 ) 2> $O/pytest_rest.time
tail -15 $O/pytest_rest.log; tail -3 $O/pytest_rest.time
