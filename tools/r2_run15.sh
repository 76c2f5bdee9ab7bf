#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_parity.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest exit $?"; tail -2 $O/pytest.log
W=C3_hall_100ktri_1Mx128_stereo
run() { # name, extra, env...
  name=$1; shift; extra=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-sub-blocks $extra > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$name.json")); s=d["stages_ms"]
    print("$name", "ms/step %.2f" % d["ms_per_step"], "image %.2f steady %.2f" % (s["image_phase"], s["steady_phase"]), d["roofline"].get("micmap"), d["roofline"]["tests_per_launch"])
except Exception as e:
    print("$name FAILED", e)
PY
}
run base "" RVB_X=1
for c in 6 8 16 24; do run cpt$c "" RVB_MICMAP_CPT=$c; done
run r131k "--rays 131072" RVB_X=1
