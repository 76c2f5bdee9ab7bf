#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown)."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    tot[r[ik]] += float(r[iv].replace(",", "")) / 1e6
    cnt[r[ik]] += 1
allms = sum(tot.values())
print("| total ms | launches | avg ms | share | kernel |\n|---:|---:|---:|---:|---|")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"| {tot[k]:.3f} | {cnt[k]} | {tot[k] / cnt[k]:.3f} | {100 * tot[k] / allms:.1f}% | `{k[:100]}` |")
print(f"\ntotal {allms:.3f} ms over {sum(cnt.values())} launches")
