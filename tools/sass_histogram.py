"""Opcode histogram per kernel of the shipped library (cuobjdump -sass), written to profiles/sass_histograms_<round>.json
together with the mnemonics that identify the hardware paths in use (DMMA, bulk TMA, mbarrier, setmaxnreg, tcgen05).
usage: python tools/sass_histogram.py r02 [lib.so]"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else 'r02'
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'dgps_with_iwvi_b200', 'lib', 'libiwvi_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
cur = None
arch = set()
for line in out.splitlines():
    m = re.match(r'\s*arch = (\S+)', line)
    if m:
        arch.add(m.group(1))
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r'\(anonymous namespace\)::', '', name).split('(')[0].replace('void ', '')
        cur = kernels.setdefault(name, collections.Counter())
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur is not None:
        cur[m.group(1)] += 1
MARK = {'DMMA': 'FP64 tensor pipe (mma.sync m8n8k4 f64)', 'UBLKCP': 'bulk TMA (cp.async.bulk)', 'SYNCS': 'mbarrier',
        'USETMAXREG': 'setmaxnreg', 'UTCQMMA': 'tcgen05.mma', 'UTCHMMA': 'tcgen05.mma', 'UTCMMA': 'tcgen05.mma',
        'LDTM': 'tcgen05.ld (TMEM)', 'STTM': 'tcgen05.st (TMEM)', 'UTCBAR': 'tcgen05.commit', 'UTCATOMSWS': 'tcgen05.alloc',
        'REDG': 'fire-and-forget red.add'}
res = {'library': os.path.relpath(lib, ROOT), 'arch': sorted(arch), 'kernels': {}}
for name, c in kernels.items():
    marks = collections.Counter()
    for op, n in c.items():
        for key in sorted(MARK, key=len, reverse=True):
            if op.split('.')[0].startswith(key):
                marks[key] += n
                break
    res['kernels'][name] = {'instructions': sum(c.values()), 'marks': dict(marks),
                            'top': dict(sorted(c.items(), key=lambda kv: -kv[1])[:25])}
path = os.path.join(ROOT, 'profiles', 'sass_histograms_%s.json' % rnd)
with open(path, 'w') as f:
    json.dump(res, f, indent=1)
tot = collections.Counter()
for k in res['kernels'].values():
    tot.update(k['marks'])
print(path, len(res['kernels']), 'kernels', dict(tot), res['arch'])
