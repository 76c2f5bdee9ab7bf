#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "band_filters or host_array" > $O/pytest_filter.log 2>&1; tail -3 $O/pytest_filter.log
echo "baseline"; timeout 300 python tools/wf_units_sweep.py 131072,1048576 0 2>&1 | grep rays | tee $O/base.log
echo "full grid (image phase)"; RVB_FULL_GRID=1 timeout 300 python tools/wf_units_sweep.py 131072,1048576 0 2>&1 | grep rays | tee $O/fullgrid.log
echo "cpt 64"; RVB_MICMAP_CPT=64 timeout 300 python tools/wf_units_sweep.py 1048576 0 2>&1 | grep rays | tee $O/cpt64.log
for sv in 20 28; do echo "service $sv"; RVB_WF_SERVICE=$sv timeout 300 python tools/wf_units_sweep.py 1048576 0 2>&1 | grep rays | tee -a $O/wf_knobs.log; done
for lm in 6 12; do echo "leaf_min $lm"; RVB_WF_LEAF_MIN=$lm timeout 300 python tools/wf_units_sweep.py 1048576 0 2>&1 | grep rays | tee -a $O/wf_knobs.log; done
