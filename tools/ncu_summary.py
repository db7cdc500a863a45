"""Summarise an .ncu-rep: per kernel key metrics + stall breakdown + hottest SASS lines.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-index ...]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']
for ki, r in enumerate(rows[2:]):
    d = dict(zip(hdr, r))
    print('=== kernel %d: %s' % (ki, d.get('Kernel Name', '')[:80]))
    for k in KEYS:
        if k in d:
            print('   %-70s %s %s' % (k, d[k], units[hdr.index(k)]))
    st = [(k, float(d[k])) for k in hdr if k.startswith('smsp__average_warp') and 'issue_stalled' in k
          and k.endswith('_per_issue_active.ratio') and d[k]]
    print('   stalls/issue:', ', '.join('%s=%.2f' % (k.split('issue_stalled_')[1].split('_per_')[0], v)
                                       for k, v in sorted(st, key=lambda kv: -kv[1])[:8]))
for arg in sys.argv[2:]:
    ki = int(arg)
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', ':::%d' % (ki + 1)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    isamp, isrc = hdr.index('# Samples'), hdr.index('Source')
    cols = [c for c in hdr if c.startswith('stall_') and '(' not in c]
    ci = [hdr.index(c) for c in cols]
    data = [r for r in rows[2:] if len(r) > max(ci) and r[isamp].replace('.', '').isdigit()]
    tot = sum(float(r[isamp]) for r in data)
    print('=== source kernel %d: %d SASS instr, %d samples' % (ki, len(data), tot))
    print('   ', ', '.join('%s=%.1f%%' % (c[6:], 100 * sum(float(r[i]) for r in data) / tot) for c, i in zip(cols, ci)
                          if sum(float(r[i]) for r in data) / tot > 0.01))
    top = sorted(range(len(data)), key=lambda k: -float(data[k][isamp]))[:25]
    for k in sorted(top):
        r = data[k]
        print('   %5d %5.1f%% %-60s %s' % (k, 100 * float(r[isamp]) / tot, r[isrc].strip()[:60],
                                          ' '.join('%s=%s' % (c[6:], r[i]) for c, i in zip(cols, ci) if float(r[i]) > 0.02 * float(r[isamp]) and float(r[i]) > 0)))
