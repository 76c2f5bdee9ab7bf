#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2m; mkdir -p $O
for cpt in 16 24 32 48; do
  echo "cells per triangle $cpt (N up to 1024)"
  RVB_MICMAP_NMAX=1024 RVB_MICMAP_CPT=$cpt timeout 300 python tools/wf_units_sweep.py 131072,1048576 0 2>&1 | grep rays | tee -a $O/sweep_cpt_fine.log
done
