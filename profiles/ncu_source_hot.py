#!/usr/bin/env python
"""Hot spots from `ncu --page source --csv` of one kernel: opcode mix by executed instructions and the
top SASS lines by stall samples.  Usage: ncu_source_hot.py file.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
mix = collections.Counter(); tot = 0; samples = []
for n, r in enumerate(rows[2:]):
    if len(r) <= iex or not r[iex].isdigit(): continue
    ex = int(r[iex] or 0); sm = int(r[ismp] or 0)
    toks = r[isrc].split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    mix[op.split('.')[0]] += ex; tot += ex
    samples.append((sm, ex, n, r[isrc].strip()))
print('total warp-instructions executed', tot)
for op, c in mix.most_common(22):
    print(f'  {op:10s} {c:10d} {100*c/tot:5.1f}%')
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print('top lines by stall samples (samples, executed, line#, sass)')
ts = sum(s[0] for s in samples)
for sm, ex, n, src in sorted(samples, reverse=True)[:top]:
    print(f'  {sm:6d} ({100*sm/max(ts,1):4.1f}%) {ex:9d} #{n:5d} {src[:90]}')
