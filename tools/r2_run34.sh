#!/bin/bash
# band filter with equal tiles + batched loads; image phase at 4 CTAs per SM
cd "$(dirname "$0")/.."
O=gpurun_out/r2iir2; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_gpu.time; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --rays 131072 --no-cpu-baseline --no-sub-blocks > $O/C3_shard131k_1gpu.json 2> $O/C3_shard131k_1gpu.err
timeout 900 python bench.py --no-cpu-baseline > $O/default_1gpu.json 2> $O/default_1gpu.err
python - <<PY
import json
for f in ["C3_shard131k_1gpu","default_1gpu"]:
    d=json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "ms/step %.2f e2e %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v,2) for k,v in d["stages_ms"].items()})
    for k,b in d.get("sub_blocks",{}).items():
        print(k, "ms/step %.2f frac %.4f" % (b["ms_per_step"], b["roofline"]["frac"]), {kk: round(v,2) for kk,v in b["stages_ms"].items()})
PY
