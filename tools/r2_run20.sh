#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
cd "$(dirname "$0")/.."
RVB_TIMING=1 timeout 600 python bench.py --workload C5_hall_64pairs_16kx128 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/c5.json 2> $O/c5.err
grep "rvb\]" $O/c5.err | grep -v "render\." | tail -24
