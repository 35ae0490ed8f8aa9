#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (profiles/sass_summary.txt): which kernels contain TMA bulk
copies (UBLKCP), mbarrier operations (SYNCS), DSMEM async stores (STAS), cluster barriers (UCGABAR), packed fp32 FMAs
(FFMA2), 128-bit loads — and that none contains tensor-core instructions (the path is HBM-bound byte work).

    python tools/sass_summary.py [path/to/libtrb200.so] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'tensor_regression_b200', 'libtrb200.so')
KEYS = ['UBLKCP', 'SYNCS', 'STAS', 'UCGABAR', 'FFMA2', 'FFMA', 'DFMA', 'LDG.E.128', 'LDS.128', 'LDS.64', 'STS', 'SHFL',
        'UTCMMA', 'HMMA', 'LDTM', 'MEMBAR', 'CCTL', 'BAR.SYNC', 'STL', 'LDL']
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
demangle = {}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        counts[cur]['_total'] += 1
        for k in KEYS:
            if op == k or op.startswith(k + '.') or (k in ('LDG.E.128', 'LDS.128', 'LDS.64') and op.startswith(k.split('.')[0]) and k.split('.', 1)[1] in op):
                counts[cur][k] += 1
names = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
print(f'# SASS mnemonic counts per kernel of {os.path.relpath(so, ROOT)} (cuobjdump -sass; FFMA counts include FFMA2)')
print('# ' + ' '.join(f'{k:>9s}' for k in ['instr'] + KEYS) + '  kernel')
for (mangled, c), nm in zip(counts.items(), names):
    nm = re.sub(r'\(.*$', '', nm)
    nm = re.sub(r'^void ', '', nm)
    print('  ' + ' '.join(f'{c[k]:9d}' for k in ['_total'] + KEYS) + '  ' + nm)
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print('# total: ' + ', '.join(f'{k}={tot[k]}' for k in KEYS if tot[k]))
print('# tensor-core instructions (UTCMMA/HMMA/LDTM): %d — none expected: the path is HBM-bound, fp32/fp64 CUDA-core FMAs keep the 1e-5 / 1e-10 tolerance' % (tot['UTCMMA'] + tot['HMMA'] + tot['LDTM']))
