#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
cd "$(dirname "$0")/.."
for W in C5_hall_64pairs_16kx128 C1_vault_3754tri_1024x64; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > $O/$W.json 2> $O/$W.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$W.json")); s=d["stages_ms"]
    print("$W", "ms/step %.2f e2e %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v,2) for k,v in s.items()}, d["bounces"])
except Exception as e:
    print("$W FAILED", e)
PY
done
