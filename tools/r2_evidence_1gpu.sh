#!/bin/bash
# Round-2 evidence on one B200: GPU tests, bench lines, ncu launch list and full captures.
# Each ncu pass runs only after the same command has exited 0 without ncu.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2ev; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_gpu.time
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/default_1gpu.json 2> $O/default_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/default_reference_cpu.json 2> $O/default_reference_cpu.err
for wl in C1_vault_3754tri_1024x64 C5_hall_64pairs_16kx128; do
  timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline > $O/${wl}_1gpu.json 2> $O/${wl}_1gpu.err
done
# C4's mesh and output model on one GPU at a reduced ray count (the full 4 Mi x 256 is the 8-GPU configuration)
timeout 900 python bench.py --workload C4_cathedral_1Mtri_4Mx256_hrtf --rays 262144 --steps 2 --no-cpu-baseline > $O/C4_cathedral_262kx256_1gpu.json 2> $O/C4_cathedral_262kx256_1gpu.err
# the shard-sized problem of the 8-GPU run (131072 rays) on one GPU: where a shard's time goes
timeout 600 python bench.py --rays 131072 --no-cpu-baseline --no-sub-blocks > $O/C3_shard131k_1gpu.json 2> $O/C3_shard131k_1gpu.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks"
timeout 600 $CMD > $O/plain_for_ncu.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
CMD4="python bench.py --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-blocks"
timeout 600 $CMD4 > $O/plain_for_ncu4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wavefront|trace_bvh_kernel|reduce_cta|iir_kernel|FlattenIn" -c 10 -o /tmp/prof_c3_kernels -f $CMD4 > $O/ncu_full_c3.log 2>&1
CMD2="python bench.py --workload C2_shoebox_12tri_1Mx128_omni --rays 262144 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD2 > $O/plain_for_ncu2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"trace_bruteforce|reduce_cta|iir_kernel|FlattenIn" -c 10 -o /tmp/prof_c2_kernels -f $CMD2 > $O/ncu_full_c2.log 2>&1
# the reports stay on the box (gpurun_out/ is capped at 64 MiB): bring back the summaries and the raw metric pages
for r in c3 c2; do
  python tools/ncu_summary.py /tmp/prof_${r}_kernels.ncu-rep > $O/ncu_${r}_summary.md 2> $O/ncu_${r}_summary.err
  ncu -i /tmp/prof_${r}_kernels.ncu-rep --page raw --csv > $O/ncu_${r}_raw.csv 2>/dev/null
done
# per-instruction stall samples of the band filter (source page, SASS)
ncu -i /tmp/prof_c3_kernels.ncu-rep --page source --csv --kernel-name-base demangled --kernel-name regex:iir_kernel > $O/ncu_iir_source.csv 2>/dev/null
ls -la $O
python - <<PY
import json
d=json.loads(open("$O/default_1gpu.json").read().strip().splitlines()[-1])
print("C3 ms/step %.2f value %.4g e2e %.4g roofline %.4f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
print("stages", {k: round(v,2) for k,v in d["stages_ms"].items()})
for k,b in d.get("sub_blocks",{}).items():
    print(k, "ms/step %.2f frac %.4f" % (b["ms_per_step"], b["roofline"]["frac"]), {kk: round(v,2) for kk,v in b["stages_ms"].items()})
PY
