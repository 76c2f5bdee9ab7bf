#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 600 tools/sort_test.bin > $O/sort_test.log 2>&1; echo "rc=$?" >> $O/sort_test.log
grep -E "FAIL|correctness|rc=|ours" $O/sort_test.log | head -40
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"onesweep_pass|digit_histogram|digit_scan" -c 8 --csv --log-file $O/ncu_times.csv tools/sort_test.bin bench > $O/ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/ncu_times.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii], r[ki][:80]),{})[r[mi]]=r[vi]
for (i,k),m in cur.items(): print(i,k,m)
PY
