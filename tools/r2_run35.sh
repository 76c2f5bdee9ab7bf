#!/bin/bash
# where the non-kernel part of a shard-sized trace goes (stage timers synchronise: the sum exceeds the untimed step)
cd "$(dirname "$0")/.."
O=gpurun_out/r2t; mkdir -p $O
RVB_TIMING=1 timeout 600 python bench.py --rays 131072 --steps 2 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/timing.json 2> $O/timing.log
tail -60 $O/timing.log
