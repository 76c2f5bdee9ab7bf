#!/bin/bash
O=gpurun_out/r2ac; mkdir -p $O
cd "$(dirname "$0")/.."
for g in "" 32 64 128; do
  echo "=== granularity arg '$g'"
  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none tools/gather_bench.bin 0 $g 2>&1 | grep -E "default|set to|mode|dram__bytes_read|gpu__time" | tail -6
done 2>&1 | tee $O/gather_granularity.log
