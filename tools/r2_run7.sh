#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 600 python -u -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c5_worker" > $O/c5_fix.log 2>&1
echo "== exit $?"; tail -c 1200 $O/c5_fix.log | tail -12
