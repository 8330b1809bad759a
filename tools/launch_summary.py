#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and share.  Usage: tools/launch_summary.py launches.csv [first_id last_id]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
agg = OrderedDict()
tot = 0.0
for r in rows:
    if not (lo <= int(r[0]) <= hi):
        continue
    name = r[4].split("(")[0].replace("void ", "").replace("satmvs::", "")[:70] + " grid=" + r[8] + " blk=" + r[7]
    ns = float(r[14])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
    tot += ns
print(f"launches {sum(a[0] for a in agg.values())}  total {tot / 1e3:.1f} us")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{ns / 1e3:10.1f} us {100 * ns / tot:5.1f}%  x{n:<4d} avg {ns / n / 1e3:8.2f} us  {name}")
