"""Opcode histogram (executed warp instructions) and top stall lines from an `ncu --page source --csv` dump.
    python scripts/ncu_opcodes.py gpurun_out/x_source.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
c = collections.Counter()
tot = 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == 'Address':
        continue
    try:
        n = int(r[ix['Instructions Executed']])
        st = int(r[ix['Warp Stall Sampling (All Samples)']])
    except ValueError:
        continue
    toks = r[ix['Source']].split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '')
    c[op.split('.')[0]] += n
    tot += n
    lines.append((st, n, r[ix['Source']].strip()))
print('executed warp instructions:', tot)
for k, v in c.most_common(topn):
    print(f'  {k:12s} {v:12d} {100 * v / tot:5.1f}%')
ssum = sum(l[0] for l in lines)
print('top stall lines (samples, share, executed, sass):')
for st, n, s in sorted(lines, reverse=True)[:topn]:
    print(f'  {st:7d} {100 * st / max(1, ssum):5.1f}% {n:10d}  {s[:100]}')
