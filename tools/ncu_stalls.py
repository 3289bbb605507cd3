#!/usr/bin/env python
"""Top stall lines of an ncu report. usage: tools/ncu_stalls.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs, cur = [], None
for r in rows:
    if r and r[0] == 'File Path':
        cur = dict(file=r[1], rows=[]); secs.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
agg = []
for s in secs:
    h = None
    for k, r in enumerate(s['rows']):
        if 'Instructions Executed' in r:
            h, st = r, k + 1
            break
    if not h:
        continue
    ln, sp = h.index('Line No'), h.index('# Samples')
    cols = {n: h.index(n) for n in ('stall_long_sb', 'stall_barrier', 'stall_short_sb', 'stall_wait', 'stall_not_selected', 'stall_math', 'stall_branch_resolving')}
    for r in s['rows'][st:]:
        if len(r) <= sp or not r[ln].strip():
            continue
        try:
            agg.append((int(float(r[sp] or 0)), {n: int(float(r[i] or 0)) for n, i in cols.items()}, s['file'].split('/')[-1], int(r[ln]), r[ln + 1].strip()[:78]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg) or 1
print('total samples', tot)
for a in sorted(agg, key=lambda x: -x[0])[:top]:
    d = a[1]
    print('%5.1f%% lsb %4d ssb %4d wait %4d nsel %4d | %s:%d %s' % (100 * a[0] / tot, d['stall_long_sb'], d['stall_short_sb'], d['stall_wait'], d['stall_not_selected'], a[2], a[3], a[4]))
