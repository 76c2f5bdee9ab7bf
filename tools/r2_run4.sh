#!/bin/bash
# round 2, GPU call 4: traversal stacks in shared memory
O=gpurun_out/r2d; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_parity.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
W=C3_hall_100ktri_1Mx128_stereo
run() { # name, extra args, env...
  name=$1; shift; extra=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline $extra > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$name.json")); s=d["stages_ms"]
    print("$name", "ms/step %.2f" % d["ms_per_step"], "image %.2f steady %.2f keys %.2f sort %.2f reduce %.2f" % (s["image_phase"], s["steady_phase"], s["keys"], s["sort"], s["reduce"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
run base "" RVB_X=1
run wf0 "" RVB_WAVEFRONT=0
run sv24 "" RVB_WF_SERVICE=24
run sv24_minb5 "" RVB_WF_SERVICE=24 RVB_WF_MINB=5
run sv24_minb7 "" RVB_WF_SERVICE=24 RVB_WF_MINB=7
run sv24_lm4 "" RVB_WF_SERVICE=24 RVB_WF_LEAF_MIN=4
run r131k "--rays 131072" RVB_X=1
run r131k_wf0 "--rays 131072" RVB_WAVEFRONT=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_bvh" -s 6 -c 2 -o $O/prof_wf python bench.py --workload $W --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
RVB_WAVEFRONT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_bvh_kernel<0" -s 3 -c 1 -o $O/prof_wf0 python bench.py --workload $W --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full0.log 2>&1
ls $O | wc -l
