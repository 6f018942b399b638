#!/usr/bin/env python
"""Per-kernel mean duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = defaultdict(list)
for r in rows[1:]:
    d[r[ki][:70]].append(float(r[vi].replace(',', '')))
for k, v in d.items():
    print(f'{k:70s} n={len(v):3d} mean={sum(v) / len(v) / 1000:8.2f} us')
