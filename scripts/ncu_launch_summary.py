#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches, total and last duration (ms)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
    agg.setdefault(r[ki], []).append(v)
for k, v in agg.items():
    print(f"{k[:70]:70s} n={len(v):5d} total={sum(v):10.3f} ms  last={v[-1]:.4f} ms")
