#!/usr/bin/env python
"""Reads gpurun_out/abncu_*.csv (tools/ab_ncu.sh): per variant and launch, duration, warp instructions, issue-active."""
import csv, glob, os, sys
for path in sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'abncu_*.csv'))):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    iid, iname, ival, imet = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    by = {}
    for r in rows[1:]:
        by.setdefault((r[iid], r[iname][:60]), {})[r[imet]] = float(r[ival].replace(',', ''))
    for (i, k), m in by.items():
        print('%-14s %s %-50s %9.1f us  %12.0f inst  issue %.1f%%' % (os.path.basename(path)[6:-4], i, k, m.get('gpu__time_duration.sum', 0) / 1e3,
              m.get('smsp__inst_executed.sum', 0), m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)))
