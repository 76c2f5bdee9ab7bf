#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest exit $?"; tail -2 $O/pytest.log
run() { # name, workload, env...
  name=$1; shift; W=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/$name.json")); s=d["stages_ms"]
    print("$name", "ms/step %.2f" % d["ms_per_step"], "image %.2f steady %.2f frac %.4f" % (s["image_phase"], s["steady_phase"], d["roofline"]["frac"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
for b in 0 640; do run c2_b$b C2_shoebox_12tri_1Mx128_omni RVB_BLOCK=$b; done
for b in 0 640; do run x_b$b X_room_4ktri_303kx64_bruteforce RVB_BLOCK=$b; done
