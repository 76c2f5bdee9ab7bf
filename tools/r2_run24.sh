#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 600 tools/sort_test.bin > $O/sort_test.log 2>&1; echo "rc=$?" >> $O/sort_test.log
grep -E "FAIL|correctness|rc=|ours" $O/sort_test.log | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass|digit_histogram" -c 6 -o $O/sort_ncu tools/sort_test.bin bench > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ncu -i $O/sort_ncu.ncu-rep --page raw --csv > $O/sort_ncu_raw.csv 2>/dev/null
ls -la $O
