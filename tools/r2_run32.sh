#!/bin/bash
# segmented wavefront kernel: parity test, units sweep, then the round's evidence run
cd "$(dirname "$0")/.."
O=gpurun_out/r2seg; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "work_units or bvh" > $O/pytest_seg.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_seg.log
timeout 600 python tools/wf_units_sweep.py > $O/sweep_minb5.log 2>&1; cat $O/sweep_minb5.log
RVB_WF_MINB=6 timeout 300 python tools/wf_units_sweep.py 131072 1,0,8 > $O/sweep_minb6.log 2>&1; cat $O/sweep_minb6.log
RVB_WF_MINB=7 timeout 300 python tools/wf_units_sweep.py 131072 1,0,8 > $O/sweep_minb7.log 2>&1; cat $O/sweep_minb7.log
bash tools/r2_evidence_1gpu.sh
