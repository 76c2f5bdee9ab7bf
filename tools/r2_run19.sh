#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_parity.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest exit $?"; tail -3 $O/pytest.log
run() { # name, workload, env...
  name=$1; shift; W=$1; shift
  env "$@" timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-sub-blocks > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$name.json")); s=d["stages_ms"]
    print("$name", "ms/step %.2f e2e %.2f frac %.4f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]), {k: round(v,2) for k,v in s.items()}, d["roofline"].get("micmap"))
except Exception as e:
    print("$name FAILED", e)
PY
}
run c5_w4 C5_hall_64pairs_16kx128 RVB_BENCH_WORKERS=4
run c5_w8 C5_hall_64pairs_16kx128 RVB_BENCH_WORKERS=8
run c5_w2 C5_hall_64pairs_16kx128 RVB_BENCH_WORKERS=2
run c3 C3_hall_100ktri_1Mx128_stereo RVB_X=1
run c2 C2_shoebox_12tri_1Mx128_omni RVB_X=1
