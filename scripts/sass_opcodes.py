"""Static SASS opcode counts per kernel of libb200gan.so (`cuobjdump -sass`), as a markdown table.
    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'gan_control_b200', 'libb200gan.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
funcs, cur = collections.OrderedDict(), None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = next(it)
        cur = re.sub(r'\(.*', '', cur).replace('b200gan::', '').replace('void ', '')
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
cols = ['UTCHMMA', 'LDTM', 'UTMALDG', 'UTCBAR', 'SYNCS', 'UTMASTG', 'HMMA', 'FADD2', 'FMUL2', 'FFMA2']
print('# Round 2 -- SASS opcode counts per kernel of `libb200gan.so` (`cuobjdump -sass`, sm_100a), final build\n')
print('Static instruction counts in the shipped library (`python scripts/sass_opcodes.py`).  `UTCHMMA` = `tcgen05.mma` (kind::f16, bf16 or\n'
      'fp16 operands), `LDTM` = `tcgen05.ld`, `UTMALDG` = TMA tensor loads (`cp.async.bulk.tensor`), `UTCBAR` = `tcgen05.commit`, `SYNCS` =\n'
      'mbarrier operations, `HMMA` = warp-level `mma.sync` (ONLY in the 1x1 RGB-side kernels `pw_small_*_mma_kernel`, where M = 16-pixel\n'
      'granularity and K <= 32 leave nothing to stage for tcgen05), `FADD2 / FMUL2 / FFMA2` = packed fp32 pairs (sm_100).  No CUTLASS /\n'
      'CuTe symbols anywhere in the library.\n')
print('| kernel | SASS instr. | ' + ' | '.join(cols) + ' |')
print('|---|---|' + '---|' * len(cols))
rows = sorted(funcs.items(), key=lambda kv: (-kv[1]['UTCHMMA'], -kv[1]['HMMA'], -kv[1]['UTMALDG'], kv[0]))
for name, c in rows:
    if sum(c.values()) < 200 and not any(c[k] for k in cols):
        continue
    print(f'| `{name}` | {sum(c.values())} | ' + ' | '.join(str(c[k]) for k in cols) + ' |')
