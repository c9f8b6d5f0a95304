#!/usr/bin/env python
"""top stall sites of one kernel from `ncu -i rep --page source --csv --print-source sass`"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
h = r[1]
c, ie = h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
rows = [(int(x[c] or 0), int(x[ie] or 0), x[1].strip(), x[0]) for x in r[2:] if len(x) > ie]
tot = sum(a for a, _, _, _ in rows)
print('total samples', tot, 'instructions executed', sum(b for _, b, _, _ in rows))
for i, (a, b, s, ad) in enumerate(rows):
    if a > tot * frac:
        print('%5d %s %6d (%4.1f%%) exec %8d  %s' % (i, ad[-5:], a, 100.0 * a / tot, b, s[:100]))
