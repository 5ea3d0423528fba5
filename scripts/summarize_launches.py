"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
    python scripts/summarize_launches.py gpurun_out/launches.csv [title]"""
import collections
import csv
import io
import re
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
text = open(path, errors='replace').read()
start = text.find('"ID"')
rows = list(csv.DictReader(io.StringIO(text[start:])))
agg = collections.OrderedDict()
total = 0.0
for r in rows:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3)
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'<.*', '<>', name) if len(name) > 70 else name
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    total += v
print(f'# {title}\n')
print(f'{sum(a[0] for a in agg.values())} launches, {total / 1e3:.2f} ms summed `gpu__time_duration` '
      f'(ncu serialises launches and runs them cold-cache: use the SHARES, not the absolute times)\n')
print('| kernel | launches | total ms | share | avg us |')
print('|---|---|---|---|---|')
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{name}` | {n} | {us / 1e3:.3f} | {100 * us / total:.1f} % | {us / n:.1f} |')
