#!/bin/bash
# new band filter + segmented wavefront (read-back fix): tests, image-phase occupancy knob, evidence
cd "$(dirname "$0")/.."
O=gpurun_out/r2iir; mkdir -p $O
timeout 300 python tools/wf_units_sweep.py 131072,1048576 0 > $O/sweep_img1.log 2>&1; cat $O/sweep_img1.log
RVB_IMG_MINB=4 timeout 300 python tools/wf_units_sweep.py 131072,1048576 0 > $O/sweep_img4.log 2>&1; cat $O/sweep_img4.log
bash tools/r2_evidence_1gpu.sh
