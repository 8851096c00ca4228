#!/usr/bin/env python
"""Opcode histogram + hottest SASS lines of one kernel from `ncu --page source --csv`.
    ncu -i X.ncu-rep --page source --csv > /tmp/src.csv ; python scripts/ncu_sass.py /tmp/src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows[:10]) if 'Source' in r)
hdr = rows[hi]
si, ei, sm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops = collections.Counter()
samples = collections.Counter()
tot = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) <= ei or not r[ei]:
        continue
    try:
        n = int(float(r[ei])); s = int(float(r[sm] or 0))
    except ValueError:
        continue
    src = r[si].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL')) and '.' in op else '')
    ops[op] += n; samples[op] += s; tot += n
    lines.append((s, n, src))
print('total warp instructions', tot)
for op, n in ops.most_common(28):
    print(f'{op:14s} {n:12d} {100.0 * n / tot:5.1f}%   samples {samples[op]}')
print('--- hottest lines by stall samples')
for s, n, src in sorted(lines, reverse=True)[:top]:
    print(f'{s:7d} {n:10d}  {src[:110]}')
