#!/bin/bash
# wavefront kernel re-tune after the 64-byte nodes
O=gpurun_out/r2o; mkdir -p $O
cd "$(dirname "$0")/.."
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
for sv in 8 12 16 20 28; do run sv$sv "" RVB_WF_SERVICE=$sv; done
for lm in 2 4 12 16; do run lm$lm "" RVB_WF_LEAF_MIN=$lm; done
for m in 4 6 7; do run minb$m "" RVB_WF_MINB=$m; done
run sv16_lm4 "" RVB_WF_SERVICE=16 RVB_WF_LEAF_MIN=4
run r262k_wf "--rays 262144" RVB_WAVEFRONT=1
run r262k_pr "--rays 262144" RVB_WAVEFRONT=0
run r262k_wf_minb7 "--rays 262144" RVB_WAVEFRONT=1 RVB_WF_MINB=7
run r131k_wf_minb5 "--rays 131072" RVB_WAVEFRONT=1 RVB_WF_MINB=5
run r131k_wf_minb4 "--rays 131072" RVB_WAVEFRONT=1 RVB_WF_MINB=4
run r524k_wf "--rays 524288" RVB_WAVEFRONT=1
run r524k_pr "--rays 524288" RVB_WAVEFRONT=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_bvh_wavefront" -s 3 -c 1 -o $O/prof_wf python bench.py --workload $W --rays 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/ncu_wf.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"iir_kernel" -s 3 -c 1 -o $O/prof_iir python bench.py --workload $W --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/ncu_iir.log 2>&1
ls $O/*.ncu-rep
