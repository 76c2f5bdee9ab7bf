#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
cd "$(dirname "$0")/.."
timeout 600 tools/sort_test.bin > $O/sort_test.log 2>&1; echo "rc=$?" >> $O/sort_test.log
grep -E "FAIL|correctness|rc=|ours" $O/sort_test.log | head -40
for v in r512 r768; do
  timeout 300 tools/sort_test_$v.bin > $O/sort_test_$v.log 2>&1; echo "rc=$?" >> $O/sort_test_$v.log
  echo "== $v"; grep -E "correctness|rc=|ours" $O/sort_test_$v.log
done
