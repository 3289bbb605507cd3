#!/usr/bin/env python
"""Save the judged summaries of an ncu report under profiles/.
usage: tools/save_profile.py rep tag "<note>" [ticks_per_launch]
With ticks_per_launch the summary is also filed in profiles/tick_kernel_ncu_summary.json under that key: bench.py prints
`roofline.traffic` only beside a launch of the shape it was measured on."""
import csv, io, json, os, subprocess, sys
rep, tag, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
key = sys.argv[4] if len(sys.argv) > 4 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, r = rows[0], rows[1], rows[2]
def get(k):
    i = hdr.index(k); v = float(r[i]); u = units[i]
    return v * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9}.get(u, 1)
summ = dict(kernel=r[hdr.index('Kernel Name')], note=note,
            duration_us=get('gpu__time_duration.sum') * 1e6,
            dram_bytes_read=get('dram__bytes_read.sum'), dram_bytes_write=get('dram__bytes_write.sum'),
            dram_bytes_per_launch=get('dram__bytes_read.sum') + get('dram__bytes_write.sum'),
            warp_instructions=get('smsp__inst_executed.sum'),
            registers_per_thread=get('launch__registers_per_thread'),
            dram_throughput_pct_of_nominal_peak=get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            issue_active_pct=get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
            achieved_occupancy_pct=get('sm__warps_active.avg.pct_of_peak_sustained_active'),
            l2_hit_pct=get('lts__t_sector_hit_rate.pct'))
if key is not None:
    path = os.path.join(root, 'profiles', 'tick_kernel_ncu_summary.json')
    try:
        table = json.load(open(path))
        if 'kernel' in table:          # (the round-1 file held one summary)
            table = {}
    except Exception:
        table = {}
    table[str(int(key))] = summ
    json.dump(table, open(path, 'w'), indent=1)
json.dump(summ, open(os.path.join(root, 'profiles', '%s_ncu_summary.json' % tag), 'w'), indent=1)
det = subprocess.run(['ncu', '-i', rep, '--page', 'details', '--csv'], capture_output=True, text=True).stdout
drows = list(csv.reader(io.StringIO(det))); dh = drows[0]
with open(os.path.join(root, 'profiles', '%s_ncu_details.txt' % tag), 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on; %s\n' % note)
    for d in drows[1:]:
        if d[dh.index('Metric Name')]:
            f.write(' | '.join(d[dh.index(k)] for k in ('Section Name', 'Metric Name', 'Metric Unit', 'Metric Value')) + '\n')
for tool, suffix in (('ncu_summary.py', 'lines'), ('ncu_stalls.py', 'stalls')):
    txt = subprocess.run([sys.executable, os.path.join(root, 'tools', tool), rep] + (['32768', '45'] if suffix == 'lines' else ['25']),
                         capture_output=True, text=True).stdout
    open(os.path.join(root, 'profiles', '%s_ncu_%s.txt' % (tag, suffix)), 'w').write(txt)
print(json.dumps(summ, indent=1))
