#!/bin/bash
O=gpurun_out/r2ab; mkdir -p $O
cd "$(dirname "$0")/.."
for wl in C2_shoebox_12tri_1Mx128_omni C3_hall_100ktri_1Mx128_stereo; do
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"onesweep_pass|digit_histogram|digit_scan|reduce_" -c 12 --csv --log-file $O/ncu_$wl.csv python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --no-sub-blocks > $O/ncu_$wl.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/ncu_$wl.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii], r[ki][:100]),{})[r[mi]]=r[vi]
for (i,k),m in list(cur.items())[:8]: print(i,k,m)
PY
done
