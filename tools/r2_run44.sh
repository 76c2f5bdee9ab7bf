#!/bin/bash
# per-instruction stall samples of the wavefront kernel (C3, one 131 k-ray shard)
cd "$(dirname "$0")/.."
O=gpurun_out/r2w; mkdir -p $O
CMD="python bench.py --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wavefront" -s 2 -c 1 -o /tmp/prof_wf -f $CMD > $O/ncu_wf.log 2>&1
ncu -i /tmp/prof_wf.ncu-rep --page source --csv > $O/ncu_wf_source.csv 2>/dev/null
ncu -i /tmp/prof_wf.ncu-rep --page source --csv --print-source cuda > $O/ncu_wf_source_cuda.csv 2>/dev/null
ls -la $O; tail -3 $O/ncu_wf.log
