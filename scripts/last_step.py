#!/usr/bin/env python
"""Print the kernels of the last update+lookup step from an ncu launch list (csv)."""
import csv, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, x in enumerate(rows) if 'op_begin' in x['Kernel Name']]
start = idx[-2]
tot = 0
for x in rows[start:]:
    k = re.sub(r'\(.*', '', x['Kernel Name'])
    k = re.sub(r'unnamed>::|void |hb::|<unnamed>::', '', k)[:64]
    v = float(x['Metric Value'].replace(',', '')) / 1e3
    tot += v
    print(f"{k:66s} grid={x['Grid Size']:>12s} {v:8.2f}us")
print('total %.1f us over %d launches' % (tot, len(rows) - start))
