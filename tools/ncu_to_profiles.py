"""Condense an `ncu --set full` report into the tracked evidence under profiles/:
    profiles/ncu_<config>_summary.json   per kernel (the longest launch of each): duration, DRAM bytes, tensor/DMMA pipe
                                         activity, registers, shared memory, stall mix  (bench.py reads dram bytes from here)
    profiles/ncu_<config>_<round>.md     the same as a table plus the hottest source lines of each kernel
usage: python tools/ncu_to_profiles.py gpurun_out/prof.ncu-rep[,second.ncu-rep,...] c3 r01"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
reps, config, rnd = sys.argv[1].split(','), sys.argv[2], sys.argv[3]
rep = reps[0]
rows = None
for rp in reps:       # several single-kernel reports are concatenated (same metric set, same columns)
    raw = subprocess.run(['ncu', '-i', rp, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    for r in rr[2:]:
        r.append(rr[1])      # units differ between reports (Kbyte / Mbyte, us / ms): each row keeps its own
        r.append(rp)
    if rows is None:
        rows = rr
    else:
        assert rr[0] == rows[0], 'reports with different metric sets'
        rows += rr[2:]
hdr, units = rows[0], rows[1]


def col(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith('.' + name) or h.endswith(name):
            return i
    return None


WANT = {
    'duration_us': 'gpu__time_duration.sum',
    'dram_read_bytes': 'dram__bytes_read.sum',
    'dram_write_bytes': 'dram__bytes_write.sum',
    'tensor_pipe_active_pct': 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'dmma_cycles_active_per_smsp': 'smsp__pipe_tensor_subpipe_dmma_cycles_active.avg',
    'sm_cycles_elapsed_max': 'sm__cycles_elapsed.max',
    'registers_per_thread': 'launch__registers_per_thread',
    'grid': 'launch__grid_size',
    'block': 'launch__block_size',
    'dyn_smem_bytes': 'launch__shared_mem_per_block_dynamic',
    'l2_hit_pct': 'lts__t_sector_hit_rate.pct',
    'issue_active_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smem_bank_conflicts': 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
}
idx = {k: col(v) for k, v in WANT.items()}
iname = hdr.index('Kernel Name')


def num(r, k):
    i = idx[k]
    if i is None or not r[i]:
        return None
    try:
        v = float(r[i].replace(',', ''))
    except ValueError:      # 'no data'
        return None
    u = r[-2][i]
    if k.endswith('_bytes') and u in ('Kbyte', 'Mbyte', 'Gbyte'):
        v *= {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    if k == 'duration_us' and u in ('ns', 'ms', 's'):
        v *= {'ns': 1e-3, 'ms': 1e3, 's': 1e6}[u]
    return v


best = {}
for ki, r in enumerate(rows[2:]):
    name = r[iname]
    short = name.split('(')[0].split('::')[-1].split('<')[0]
    d = {k: num(r, k) for k in WANT}
    d['kernel_index'] = ki if len(reps) == 1 else 0
    d['report'] = os.path.basename(r[-1])
    d['full_name'] = name[:120]
    st = [(h.split('issue_stalled_')[1].split('_per_')[0], float(r[i])) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and r[i]]
    d['stalls_per_issue'] = {k: round(v, 2) for k, v in sorted(st, key=lambda kv: -kv[1])[:6]}
    if d['dram_read_bytes'] is not None and d['dram_write_bytes'] is not None:
        d['dram_bytes_per_launch'] = d['dram_read_bytes'] + d['dram_write_bytes']
    if d['dmma_cycles_active_per_smsp'] and d['sm_cycles_elapsed_max']:
        d['dmma_pipe_busy_frac'] = d['dmma_cycles_active_per_smsp'] / d['sm_cycles_elapsed_max']
    if short not in best or (d['duration_us'] or 0) > (best[short]['duration_us'] or 0):
        best[short] = d

out = {'config': config, 'round': rnd, 'report': ','.join(os.path.basename(x) for x in reps),
       'note': 'per-launch values from one `ncu --set full --clock-control none` capture (cold caches, serialised '
               'launches); the longest launch of each kernel = an inner GP layer', 'kernels': best}
os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
with open(os.path.join(ROOT, 'profiles', 'ncu_%s_summary.json' % config), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)

md = ['# ncu --set full, config %s, %s (%s)\n' % (config, rnd, ', '.join(os.path.basename(x) for x in reps)),
      '| kernel | us | DMMA pipe busy | tensor pipe active % | DRAM read MB | DRAM write MB | regs | dyn smem KB | L2 hit % |',
      '|---|---|---|---|---|---|---|---|---|']
for k, d in sorted(best.items()):
    f2 = lambda v, s=1.0: '-' if v is None else '%.1f' % (v * s)
    md.append('| %s | %s | %s | %s | %s | %s | %s | %s | %s |' % (
        k, f2(d['duration_us']), '-' if 'dmma_pipe_busy_frac' not in d else '%.3f' % d['dmma_pipe_busy_frac'],
        f2(d['tensor_pipe_active_pct']), f2(d['dram_read_bytes'], 1e-6), f2(d['dram_write_bytes'], 1e-6),
        f2(d['registers_per_thread']), f2(d['dyn_smem_bytes']), f2(d['l2_hit_pct'])))
md.append('')
for k, d in sorted(best.items()):
    md.append('## %s (launch %d): stalls per issue %s\n' % (k, d['kernel_index'], d['stalls_per_issue']))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'),
                        os.path.join(os.path.dirname(rep), d['report']), str(d['kernel_index']), k],
                       capture_output=True, text=True)
    md.append('```\n' + (r.stdout.strip() or r.stderr.strip())[:6000] + '\n```\n')
with open(os.path.join(ROOT, 'profiles', 'ncu_%s_%s.md' % (config, rnd)), 'w') as f:
    f.write('\n'.join(md))
print('\n'.join(md[:12]))
