#!/bin/bash
# End-of-round multi-GPU bench lines (one rank per GPU, torchrun).
N=${1:-8}
O=gpurun_out/r1d
mkdir -p $O
shift
for wl in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --workload $wl 2> $O/${wl}_${N}gpu.err | tail -1 > $O/${wl}_${N}gpu.json
  python -c "
import json,sys
d=json.load(open('$O/${wl}_${N}gpu.json')); print('$wl', d['n_gpus'], 'ms %.2f' % d['ms_per_step'], 'e2e %.2f' % d['e2e']['ms_per_step'], 'value %.3g' % d['value'], d['config'].get('shard_balance'))" || tail -5 $O/${wl}_${N}gpu.err
done
