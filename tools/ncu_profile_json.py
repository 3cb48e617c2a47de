#!/usr/bin/env python
"""Figures of one kernel launch of an ncu report as JSON (for bench.py's roofline block and profiles/):
usage: ncu_profile_json.py report.ncu-rep kernel-regex raw_bytes_per_launch > profiles/xxx.json"""
import csv, json, re, subprocess, sys
rep, pat, raw = sys.argv[1], sys.argv[2], float(sys.argv[3])
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if re.search(pat, d.get('Kernel Name', '')):
        f = lambda k: float(d[k].replace(',', '')) if d.get(k) not in (None, '') else None
        units = dict(zip(hdr, rows[1]))
        def bytes_of(k):
            v, u = f(k), units.get(k, '')
            return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
        dram = bytes_of('dram__bytes_read.sum') + bytes_of('dram__bytes_write.sum')
        out = {
            'report': rep.split('/')[-1], 'kernel': d['Kernel Name'].split('(')[0], 'raw_bytes_per_launch': raw,
            'duration_ms_under_ncu': f('gpu__time_duration.sum') / (1e6 if units.get('gpu__time_duration.sum') == 'ns' else 1e3 if units.get('gpu__time_duration.sum') == 'us' else 1),
            'dram_bytes_per_raw_byte': dram / raw,
            'warp_inst_per_raw_byte': f('smsp__inst_executed.sum') / raw,
            'issue_slot_frac': f('smsp__issue_active.avg.pct_of_peak_sustained_active') / 100,
            'alu_pipe_frac': f('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active') / 100,
            'lsu_pipe_frac': f('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active') / 100,
            'active_lanes_per_inst': f('smsp__thread_inst_executed_per_inst_executed.ratio'),
            'occupancy_frac': f('sm__warps_active.avg.pct_of_peak_sustained_active') / 100,
            'registers_per_thread': f('launch__registers_per_thread'),
        }
        print(json.dumps(out, indent=1))
        break
