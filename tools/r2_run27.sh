#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
cd "$(dirname "$0")/.."
for v in r1024_b1 r1024_b2 r1024_b4; do
  echo "== $v"
  timeout 300 tools/sort_test_$v.bin quick > $O/st_$v.log 2>&1; grep -E "correctness" $O/st_$v.log
  timeout 300 tools/sort_test_$v.bin bench 2>&1 | tee -a $O/st_$v.log | grep ours
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"onesweep_pass|digit_histogram" -c 6 --csv --log-file $O/ncu_times_b1.csv tools/sort_test_r1024_b1.bin bench > /dev/null 2>&1
grep -E "gpu__time_duration|inst_executed" $O/ncu_times_b1.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-160
