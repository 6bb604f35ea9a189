#!/usr/bin/env python
"""Groups the listing produced by ncu_sass_regions.py ... list into regions of equal execution count."""
import re, sys
from collections import Counter
lines = open(sys.argv[1]).read().splitlines()[1:]
regs = []
for l in lines:
    m = re.match(r'\s*(\d+)\s+([\d.]+)M thr=\s*([\d.]+) smp=\s*(\d+) cum=\s*([\d.]+)\s+(.*)', l)
    if not m: continue
    i, inst, thr, smp, src = int(m[1]), float(m[2]), float(m[3]), int(m[4]), m[6]
    op = src.split()[0] if src else ''
    if op.startswith('@'): op = src.split()[1] if len(src.split()) > 1 else op
    if regs and abs(regs[-1]['inst'] - inst) <= 0.05 * max(inst, regs[-1]['inst']) + 0.005:
        r = regs[-1]; r['n'] += 1; r['tot'] += inst; r['thrsum'] += thr * inst; r['smp'] += smp; r['last'] = i; r['ops'].append(op)
    else:
        regs.append(dict(first=i, last=i, n=1, inst=inst, tot=inst, thrsum=thr * inst, smp=smp, ops=[op]))
T = sum(r['tot'] for r in regs); S = sum(r['smp'] for r in regs)
for r in regs:
    if r['tot'] / T < 0.004: continue
    c = Counter(o.split('.')[0] for o in r['ops']).most_common(6)
    print(f"[{r['first']:4d}-{r['last']:4d}] n={r['n']:3d} exec={r['inst']:6.2f}M tot={r['tot']:7.1f}M ({r['tot']/T:5.1%}) thr={r['thrsum']/max(r['tot'],1e-9):5.1f} samples={r['smp']/S:5.1%}  {c}")
