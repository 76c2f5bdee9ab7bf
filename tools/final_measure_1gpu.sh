#!/bin/bash
# End-of-round measurements on one B200: tests, bench lines, ncu launch list and full captures.
# Each ncu pass runs only after the same command has exited 0 without ncu.
set -u
O=gpurun_out/r1d
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.log
python bench.py > $O/c2_1gpu.json 2> $O/c2_1gpu.err
python bench.py --impl reference > $O/c2_reference_cpu.json 2> $O/c2_reference_cpu.err
for wl in X_room_4ktri_303kx64_bruteforce C1_vault_like_room_1024x64 C3_hall_100ktri_1Mx128_stereo C5_hall_64pairs_16kx128; do
  python bench.py --workload $wl --steps 3 > $O/${wl}_1gpu.json 2> $O/${wl}_1gpu.err
done
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $O/plain_for_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
CMD2="python bench.py --rays 262144 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD2 > $O/plain_for_ncu2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"trace_bruteforce|reduce_cta|iir_kernel|keys_kernel|bounds_kernel" -s 10 -c 6 -o $O/prof_c2_kernels -f $CMD2 > $O/ncu_full.log 2>&1
CMD3="python bench.py --workload X_room_4ktri_303kx64_bruteforce --rays 75776 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD3 > $O/plain_for_ncu3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"trace_bruteforce" -s 6 -c 2 -o $O/prof_x_kernels -f $CMD3 > $O/ncu_full_x.log 2>&1
CMD4="python bench.py --workload C3_hall_100ktri_1Mx128_stereo --rays 131072 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD4 > $O/plain_for_ncu4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"trace_bvh|reduce_" -s 6 -c 3 -o $O/prof_c3_kernels -f $CMD4 > $O/ncu_full_c3.log 2>&1
cat $O/pytest_gpu.log
tail -c 600 $O/c2_1gpu.json
