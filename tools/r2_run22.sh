#!/bin/bash
# wavefront grid / occupancy sweep at shard-sized ray counts (C3 hall): stages_ms of each run
O=gpurun_out/r2w; mkdir -p $O
cd "$(dirname "$0")/.."
run() { # name rays env...
  name=$1; rays=$2; shift 2
  env "$@" timeout 300 python bench.py --rays $rays --steps 3 --warmup 2 --no-cpu-baseline --no-sub-blocks > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$name.json")); s=d["stages_ms"]
    print("%-28s ms/step %.2f image %.2f steady %.2f keys %.2f sort %.2f reduce %.2f filter %.2f finish %.2f" % ("$name", d["ms_per_step"], s["image_phase"], s["steady_phase"], s["keys"], s["sort"], s["reduce"], s["filter"], s["finish"]))
except Exception as e: print("$name", "failed", e)
PY
}
for rays in 131072 262144; do
  run r${rays}_default $rays A=1
  for g in 444 512 592 666; do run r${rays}_g$g $rays RVB_WF_GRID=$g; done
  run r${rays}_minb4 $rays RVB_WF_MINB=4
  run r${rays}_minb6 $rays RVB_WF_MINB=6
  run r${rays}_minb7 $rays RVB_WF_MINB=7
done
run r524288_default 524288 A=1
run r524288_minb6 524288 RVB_WF_MINB=6
