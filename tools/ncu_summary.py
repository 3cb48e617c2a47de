#!/usr/bin/env python
"""Summarise .ncu-rep files (raw page) as markdown rows: one line per profiled kernel launch.
usage: ncu_summary.py report.ncu-rep [...]"""
import csv, subprocess, sys
KEYS = [('gpu__time_duration.sum', 'ms'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__registers_per_thread', 'regs'), ('launch__shared_mem_per_block_dynamic', 'dyn smem'),
        ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('smsp__inst_executed.sum', 'warp inst'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'st barrier'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'st short_sb'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'st long_sb'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'st wait')]
print('| report | kernel | ' + ' | '.join(k[1] for k in KEYS) + ' |')
print('|---|---|' + '---|' * len(KEYS))
for rep in sys.argv[1:]:
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d.get('Kernel Name', '?').split('(')[0][-40:]
        vals = []
        for k, _ in KEYS:
            v = d.get(k, '')
            try:
                f = float(v); v = ('%.3g' % f) + (' ' + u.get(k, '') if u.get(k, '') not in ('', '%') else '')
            except ValueError:
                pass
            vals.append(v)
        print('| %s | %s | %s |' % (rep.split('/')[-1], name, ' | '.join(vals)))
