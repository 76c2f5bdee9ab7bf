#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass" -c 2 -o $O/pass_ncu tools/sort_test.bin bench > $O/ncu2.log 2>&1
tail -2 $O/ncu2.log
ncu -i $O/pass_ncu.ncu-rep --page raw --csv > $O/pass_ncu_raw.csv 2>/dev/null
ncu -i $O/pass_ncu.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > $O/pass1_source.csv 2>/dev/null
ls -la $O
