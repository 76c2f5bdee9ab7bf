#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c2; mkdir -p $O
for cpt in 16 24 32 48; do echo "cpt $cpt"; RVB_MICMAP_CPT=$cpt timeout 300 python tools/wf_units_sweep.py 1048576 0 2>&1 | grep rays | tee -a $O/cpt_1m.log; done
for cpt in 6 10 16 24; do echo "cpt $cpt"; RVB_MICMAP_CPT=$cpt timeout 300 python tools/wf_units_sweep.py 131072 0 2>&1 | grep rays | tee -a $O/cpt_131k.log; done
