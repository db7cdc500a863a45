"""Attribute the PC samples of one kernel in an .ncu-rep to lines of the top-level .cu file.
Joins `ncu --page source --csv` (per-SASS-instruction samples and stall reasons) with `nvdisasm -gi` of the cubin
extracted from the library (same instruction order), using the OUTERMOST inlined-at location of each instruction.

usage: python tools/ncu_lines.py report.ncu-rep <kernel-index> <kernel-name-substring> [lib.so] [min-pct]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, ki, pat = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         'dgps_with_iwvi_b200', 'lib', 'libiwvi_b200.so')
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.7

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', ':::%d' % (ki + 1)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isamp, isrc = hdr.index('# Samples'), hdr.index('Source')
cols = [c for c in hdr if c.startswith('stall_') and '(' not in c]
ci = [hdr.index(c) for c in cols]
data = [r for r in rows[2:] if len(r) > max(ci) and r[isamp].replace('.', '').isdigit()]

tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines_of = None
for cub in sorted(glob.glob(os.path.join(tmp, '*.cubin'))):
    out = subprocess.run(['nvdisasm', '-gi', '-c', cub], capture_output=True, text=True).stdout
    cur_fn, loc, instrs = None, None, collections.defaultdict(list)
    for ln in out.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))   # last one before the instruction = outermost
            continue
        if re.match(r'\s*/\*[0-9a-f]+\*/', ln) and cur_fn:
            instrs[cur_fn].append(loc)
    for fn, locs in instrs.items():
        if pat in fn and len(data) in (len(locs), 2 * len(locs)):   # ncu lists the SASS twice when source is imported
            lines_of = locs
            data = data[:len(locs)]
if lines_of is None:
    sys.exit('no function matching %r with %d instructions found in %s' % (pat, len(data), lib))

tot = sum(float(r[isamp]) for r in data)
agg = collections.defaultdict(lambda: [0.0, collections.Counter(), 0, 0])
for r, loc in zip(data, lines_of):
    a = agg[loc]
    a[0] += float(r[isamp])
    for c, i in zip(cols, ci):
        a[1][c[6:]] += float(r[i])
    a[2] += 1
    a[3] += 'DMMA' in r[isrc]
print('kernel %d (%s): %d SASS instr, %d samples' % (ki, pat, len(data), tot))
srcfile = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if a[0] / tot * 100 < minpct:
        continue
    if f not in srcfile:
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), '..', 'csrc', f)
        srcfile[f] = open(path).read().splitlines() if os.path.exists(path) else []
    text = srcfile[f][l - 1].strip()[:70] if l - 1 < len(srcfile[f]) else ''
    top = ' '.join('%s=%.0f%%' % (k, 100 * v / a[0]) for k, v in a[1].most_common(4) if v > 0.08 * a[0])
    print('%-16s %4d %5.1f%% (%4d instr, %4d dmma) %-70s | %s' % (f, l, 100 * a[0] / tot, a[2], a[3], text, top))
