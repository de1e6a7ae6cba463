#!/usr/bin/env python
"""profiles/sass_evidence_r2.txt: per kernel of the built .so the static instruction count and the counts of the
mnemonics that prove the Blackwell paths.  usage: python tools/sass_evidence.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'cddmsl_b200', 'lib', 'libcddmsl_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout.splitlines()
names = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'REDG', 'ATOMS', 'FFMA2',
         'FMUL2']
cur, data = None, collections.OrderedDict()
for l in out:
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1)
        data[cur] = collections.Counter()
        continue
    if cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        data[cur]['instructions'] += 1
        for n in names:
            if re.search(r'\b' + n + r'\b', l):
                data[cur][n] += 1
hdr = '''SASS of cddmsl_b200/lib/libcddmsl_b200.so (cuobjdump -sass, sm_100a), round-2 final build (tools/sass_evidence.py).
Per kernel: static instruction count and the counts of the mnemonics that prove the Blackwell paths
(UTCHMMA/UTCBAR = tcgen05.mma/commit, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = 1-D bulk copies,
LDGSTS = cp.async, SYNCS = mbarrier, REDG = red.global, ATOMS = shared-memory atomics, FFMA2/FMUL2 = packed fp32x2).

'''
with open(os.path.join(ROOT, 'profiles', 'sass_evidence_r2.txt'), 'w') as f:
    f.write(hdr)
    for k, c in data.items():
        f.write(k + '\n    instructions %d  ' % c['instructions'] + '  '.join(f'{n}={c[n]}' for n in names if c[n]) + '\n')
print(len(data), "kernels")
