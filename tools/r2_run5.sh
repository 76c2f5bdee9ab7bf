#!/bin/bash
# round 2, GPU call 5: new bench.py (C3 default + sub-blocks) and the whole GPU test suite
O=gpurun_out/r2e; mkdir -p $O
cd "$(dirname "$0")/.."
( time timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
tail -3 $O/bench_default.time
python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print("C3 ms/step %.2f value %.3g e2e %.3g roofline %.4f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
print("stages", {k: round(v,2) for k,v in d["stages_ms"].items()}, "bounces", d["bounces"])
for k,b in d.get("sub_blocks",{}).items():
    print(k, "ms/step %.2f frac %.4f" % (b["ms_per_step"], b["roofline"]["frac"]), {kk: round(v,2) for kk,v in b["stages_ms"].items()})
print("cpu", d.get("cpu_baseline"))
PY
( time timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1 ) 2> $O/pytest_all.time
tail -5 $O/pytest_all.log; tail -3 $O/pytest_all.time
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
cat $O/bench_reference.json | head -c 600
