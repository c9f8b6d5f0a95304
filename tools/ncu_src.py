#!/usr/bin/env python
"""Stall reasons, opcode mix and top stall sites per kernel from
`ncu -i rep --page source --csv --print-source sass > f.csv` (handles several kernels per file)."""
import csv, sys, collections
r = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.015
secs, cur = [], None
for x in r:
    if x and x[0] == 'Kernel Name':
        cur = {'name': x[1], 'rows': []}; secs.append(cur)
    elif x and x[0] == 'Address': cur['h'] = x
    elif cur is not None and 'h' in cur and len(x) > 10: cur['rows'].append(x)
for s in secs:
    h = s['h']; c = h.index('Warp Stall Sampling (All Samples)'); ie = h.index('Instructions Executed')
    rows = s['rows']
    tot = sum(int(x[c] or 0) for x in rows) or 1; totie = sum(int(x[ie] or 0) for x in rows) or 1
    print('==', s['name'][:70], 'samples', tot, 'warp-inst', totie)
    st = [(k, sum(int(x[h.index(k)] or 0) for x in rows) / tot) for k in h if k.startswith('stall_') and 'Not Issued' not in k]
    print('  ' + ' '.join('%s=%.2f' % (k[6:], v) for k, v in sorted(st, key=lambda kv: -kv[1]) if v > 0.01))
    mix, smp = collections.Counter(), collections.Counter()
    for x in rows:
        op = x[1].strip().split()
        if op and op[0].startswith('@'): op = op[1:]
        o = op[0].split('.')[0] if op else '?'
        mix[o] += int(x[ie] or 0); smp[o] += int(x[c] or 0)
    print('  mix: ' + ' '.join('%s %.1f/%.1f' % (o, 100 * v / totie, 100 * smp[o] / tot) for o, v in mix.most_common(22)))
    for i, x in enumerate(rows):
        a = int(x[c] or 0)
        if a > tot * frac:
            print('  %5d %s %6d (%4.1f%%) exec %9d  %s' % (i, x[0][-5:], a, 100.0 * a / tot, int(x[ie] or 0), x[1].strip()[:90]))
