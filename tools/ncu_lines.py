#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line, kernel by kernel: share of executed instructions and of stall samples.
usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kernels = []   # (name, items)
cur_file = ''; hdr = None
for r in rows:
    if r and r[0] == 'Function Name':
        if not kernels or kernels[-1][1]: kernels.append((r[1], []))
        else: kernels[-1] = (r[1], [])
        continue
    if len(r) == 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if r and r[0] == 'Address': hdr = None; continue
    if hdr and len(r) > 8 and r[0].isdigit() and kernels:
        iI = hdr.index('Instructions Executed'); iS = hdr.index('Warp Stall Sampling (All Samples)')
        try: kernels[-1][1].append((cur_file, int(r[0]), r[1].strip(), int(r[iI]), int(r[iS])))
        except ValueError: pass
for name, items in kernels:
    if not items: continue
    ti = sum(x[3] for x in items) or 1; ts = sum(x[4] for x in items) or 1
    print('==== %s\n     total warp-instructions %d, stall samples %d' % (name[:110], ti, ts))
    for f, ln, src, ni, ns in items:
        if 100 * ni / ti >= thr or 100 * ns / ts >= thr:
            print('%5.1f%% inst %5.1f%% stall  %s:%-4d %s' % (100 * ni / ti, 100 * ns / ts, f, ln, src[:95]))
