#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
cd "$(dirname "$0")/.."
RVB_WF_RAYS=2 timeout 900 python -m pytest tests/test_gpu_bvh.py -x -q -m gpu > $O/pytest_wf2.log 2>&1
echo "pytest wf2 exit $?"; tail -3 $O/pytest_wf2.log
timeout 900 python -m pytest tests/test_gpu_bvh.py -x -q -m gpu > $O/pytest_wf1.log 2>&1
echo "pytest wf1 exit $?"; tail -2 $O/pytest_wf1.log
W=C3_hall_100ktri_1Mx128_stereo
run() { # name, extra, env...
  name=$1; shift; extra=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-sub-blocks $extra > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$name.json")); s=d["stages_ms"]
    print("$name", "ms/step %.2f" % d["ms_per_step"], "image %.2f steady %.2f" % (s["image_phase"], s["steady_phase"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
run wf1 "" RVB_X=1
run wf2 "" RVB_WF_RAYS=2
run wf2_minb3 "" RVB_WF_RAYS=2 RVB_WF_MINB=3
run wf2_minb5 "" RVB_WF_RAYS=2 RVB_WF_MINB=5
run wf2_sv16 "" RVB_WF_RAYS=2 RVB_WF_SERVICE=16
run wf2_sv28 "" RVB_WF_RAYS=2 RVB_WF_SERVICE=28
run wf2_131k "--rays 131072" RVB_WF_RAYS=2
run wf1_131k "--rays 131072" RVB_X=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_bvh_wavefront" -s 3 -c 1 -o $O/prof_wf2 python bench.py --workload $W --rays 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/ncu_wf2.log 2>&1
