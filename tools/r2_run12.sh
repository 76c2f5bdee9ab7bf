#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest exit $?"; tail -3 $O/pytest.log
( time timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print("C3 ms/step %.2f value %.3g e2e %.3g roofline %.4f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
print("stages", {k: round(v,2) for k,v in d["stages_ms"].items()})
for k,b in d.get("sub_blocks",{}).items():
    print(k, "ms/step %.2f frac %.4f" % (b["ms_per_step"], b["roofline"]["frac"]), {kk: round(v,2) for kk,v in b["stages_ms"].items()})
PY
