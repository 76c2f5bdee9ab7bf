#!/bin/bash
O=gpurun_out/r2aa; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 600 tools/sort_test.bin > $O/sort_test.log 2>&1; echo "rc=$?" >> $O/sort_test.log
grep -E "FAIL|correctness|rc=|ours" $O/sort_test.log | head
( time timeout 2400 python -m pytest tests -q -m gpu -x > $O/pytest_all.log 2>&1 ) 2> $O/pytest_all.time
tail -6 $O/pytest_all.log
( time timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print("C3 ms/step %.2f value %.4g e2e %.4g roofline %.4f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
print("stages", {k: round(v,2) for k,v in d["stages_ms"].items()})
for k,b in d.get("sub_blocks",{}).items():
    print(k, "ms/step %.2f frac %.4f" % (b["ms_per_step"], b["roofline"]["frac"]), {kk: round(v,2) for kk,v in b["stages_ms"].items()})
PY
RVB_FLATTEN_SORT=cub timeout 600 python bench.py --no-cpu-baseline --no-sub-blocks > $O/bench_cub.json 2> $O/bench_cub.err
python - <<PY
import json
d=json.load(open("$O/bench_cub.json"))
print("CUB path: C3 ms/step %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["stages_ms"].items()})
PY
