#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest exit $?"; tail -3 $O/pytest.log
for W in C5_hall_64pairs_16kx128 C3_hall_100ktri_1Mx128_stereo C2_shoebox_12tri_1Mx128_omni; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/$W.json 2> $O/$W.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$W.json")); s=d["stages_ms"]
    print("$W", "ms/step %.2f e2e %.2f frac %.4f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]), {k: round(v,2) for k,v in s.items()})
except Exception as e:
    print("$W FAILED", e)
PY
done
tail -3 $O/C5_hall_64pairs_16kx128.err
