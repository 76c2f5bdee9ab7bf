#!/bin/bash
# quantised 64-byte BVH nodes: parity + C3 variants; C2 re-measure with the fast f2d reduce
O=gpurun_out/r2k; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest exit $?"; tail -3 $O/pytest.log
run() { # name, workload, extra, env...
  name=$1; shift; W=$1; shift; extra=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline $extra > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$name.json")); s=d["stages_ms"]
    print("$name", "ms/step %.2f" % d["ms_per_step"], "image %.2f steady %.2f keys %.2f sort %.2f reduce %.2f frac %.4f" % (s["image_phase"], s["steady_phase"], s["keys"], s["sort"], s["reduce"], d["roofline"]["frac"]), d["roofline"].get("box_tests_per_launch"))
except Exception as e:
    print("$name FAILED", e)
PY
}
C3=C3_hall_100ktri_1Mx128_stereo
run c3_minb7 $C3 "" RVB_BVH_MINB=7
run c3_minb6 $C3 "" RVB_BVH_MINB=6
run c3_minb5 $C3 "" RVB_BVH_MINB=5
run c3_wf $C3 "" RVB_WAVEFRONT=1
run c3_wf_minb6 $C3 "" RVB_WAVEFRONT=1 RVB_WF_MINB=6
run c3_131k $C3 "--rays 131072" RVB_X=1
run c3_131k_wf $C3 "--rays 131072" RVB_WAVEFRONT=1
run c2 C2_shoebox_12tri_1Mx128_omni "" RVB_X=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_bruteforce -s 7 -c 1 -o $O/prof_c2 python bench.py --workload C2_shoebox_12tri_1Mx128_omni --rays 262144 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_bvh" -s 7 -c 1 -o $O/prof_c3 python bench.py --workload $C3 --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_c3.log 2>&1
ls $O/*.ncu-rep
