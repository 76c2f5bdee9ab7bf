#!/bin/bash
# 2-GPU: sharded bench with parity
O=gpurun_out/r2h; mkdir -p $O
cd "$(dirname "$0")/.."
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
echo "exit $?"; tail -5 $O/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("$O/bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N=%d ms/step %.2f value %.3g e2e %.3g" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
print("stages", {k: round(v,2) for k,v in d["stages_ms"].items()})
print("parity", d.get("sharded_parity"))
PY
