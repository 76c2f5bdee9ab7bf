#!/bin/bash
# round 2, GPU call 1: BVH parity + wavefront / micmap A/B on C3
mkdir -p gpurun_out/r2a
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2a/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
W=C3_hall_100ktri_1Mx128_stereo
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/$name.json 2> gpurun_out/r2a/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2a/$name.json")); s=d["stages_ms"]
    print("$name", "ms/step %.2f" % d["ms_per_step"], "image %.2f steady %.2f keys %.2f sort %.2f reduce %.2f" % (s["image_phase"], s["steady_phase"], s["keys"], s["sort"], s["reduce"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
run base_wf1_mm1 RVB_X=1
run wf0_mm1 RVB_WAVEFRONT=0
run wf0_mm0 RVB_WAVEFRONT=0 RVB_MICMAP=0
run wf1_mm0 RVB_MICMAP=0
for m in 4 5 7 8; do run wf1_minb$m RVB_WF_MINB=$m; done
for c in 2 4 16; do run wf1_cpt$c RVB_MICMAP_CPT=$c; done
# launch list + one full capture of the wavefront kernel at an eighth of the rays
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a/launches.csv python bench.py --workload $W --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_bvh" -s 6 -c 2 -o gpurun_out/r2a/prof_bvh python bench.py --workload $W --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/ncu_full.log 2>&1
ls -la gpurun_out/r2a
