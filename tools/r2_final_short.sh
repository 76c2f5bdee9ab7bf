#!/bin/bash
# last evidence pass of the round (reduced: what changed since tools/r2_evidence_1gpu.sh last ran is the BVH path only)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2fin; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_gpu.time
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/default_1gpu.json 2> $O/default_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/default_reference_cpu.json 2> $O/default_reference_cpu.err
timeout 600 python bench.py --rays 131072 --no-cpu-baseline --no-sub-blocks > $O/C3_shard131k_1gpu.json 2> $O/C3_shard131k_1gpu.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks"
timeout 600 $CMD > $O/plain_for_ncu.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
CMD4="python bench.py --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks"
timeout 600 $CMD4 > $O/plain_for_ncu4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wavefront|trace_bvh_kernel|reduce_cta|iir_kernel|FlattenIn" -c 5 -o /tmp/prof_c3_kernels -f $CMD4 > $O/ncu_full_c3.log 2>&1
python tools/ncu_summary.py /tmp/prof_c3_kernels.ncu-rep > $O/ncu_c3_summary.md 2> $O/ncu_c3_summary.err
python - <<PY
import json
d=json.loads(open("$O/default_1gpu.json").read().strip().splitlines()[-1])
print("C3 ms/step %.2f value %.4g e2e %.4g roofline %.4f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
print("stages", {k: round(v,2) for k,v in d["stages_ms"].items()})
for k,b in d.get("sub_blocks",{}).items():
    print(k, "ms/step %.2f frac %.4f" % (b["ms_per_step"], b["roofline"]["frac"]), {kk: round(v,2) for kk,v in b["stages_ms"].items()})
PY
