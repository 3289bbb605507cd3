#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU): headline metrics + instructions per CUDA source line.
usage: tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [n_warps_per_launch] [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n_warps = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40


def run(*args):
    return subprocess.run(['ncu', '-i', rep] + list(args), capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(run('--page', 'raw', '--csv'))))
hdr, units, r = rows[0], rows[1], rows[2]
for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
          'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
          'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
          'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'):
    if k in hdr:
        i = hdr.index(k)
        print('%-90s %-10s %s' % (k, units[i], r[i]))
i = hdr.index('smsp__inst_executed.sum')
print('warp-instructions per warp-tick: %.0f' % (float(r[i]) / n_warps))

rows = list(csv.reader(io.StringIO(run('--page', 'source', '--csv', '--print-source', 'cuda,sass'))))
secs, cur = [], None
for row in rows:
    if row and row[0] == 'File Path':
        cur = dict(file=row[1], rows=[])
        secs.append(cur)
    elif cur is not None:
        cur['rows'].append(row)
agg = []
for s in secs:
    h = None
    for k, row in enumerate(s['rows']):
        if 'Instructions Executed' in row:
            h, start = row, k + 1
            break
    if h is None:
        continue
    ie, ln, sp = h.index('Instructions Executed'), h.index('Line No'), h.index('# Samples')
    for row in s['rows'][start:]:
        if len(row) <= ie or not row[ln].strip():
            continue
        try:
            n, smp = int(float(row[ie])), int(float(row[sp] or 0))
        except ValueError:
            continue
        if n > 0:
            agg.append((n, smp, s['file'].split('/')[-1], int(row[ln]), row[ln + 1].strip()[:90]))
tot_s = sum(a[1] for a in agg) or 1
print('\ninstr/warp-tick  stall-samples%  file:line  source')
for n, smp, f, l, t in sorted(agg, reverse=True)[:top]:
    print('%7.1f %6.1f%%  %s:%d  %s' % (n / n_warps, 100.0 * smp / tot_s, f, l, t))
