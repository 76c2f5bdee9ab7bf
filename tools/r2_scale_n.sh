#!/bin/bash
# one point of the strong-scaling curve of the default workload (C3) + the sharded step's timeline
# usage: tools/r2_scale_n.sh N
cd "$(dirname "$0")/.."
N=${1:-8}
O=gpurun_out/r2scale; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=%d ms/step %.2f value %.4g e2e %.4g" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v,2) for k,v in d["stages_ms"].items()}, d.get("sharded_parity"))
except Exception as e:
    print("N=$N FAILED", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/shard_timeline.py > $O/timeline_${N}gpu.log 2>&1
grep "it 5" $O/timeline_${N}gpu.log | sort | head -8
tail -3 $O/bench_${N}gpu.err
