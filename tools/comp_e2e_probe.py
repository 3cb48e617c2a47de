#!/usr/bin/env python
"""Development probe: end-to-end compress (pinned host -> pinned host) versus host_batch_bytes.
usage: comp_e2e_probe.py [n_chunks] [MiB ...]"""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, synth
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 600
ns, nc = 30000, 385
cd = _native.default_codec(0)
base = [np.ascontiguousarray(synth.ap_chunk(ns, nc, seed=100 + i)) for i in range(8)]
cb = ns * nc * 2
h_raw = torch.empty(n_chunks * cb, dtype=torch.uint8, pin_memory=True)
for i in range(n_chunks): h_raw.numpy()[i * cb:(i + 1) * cb] = base[i % 8].reshape(-1).view(np.uint8)
rows = np.arange(n_chunks + 1, dtype=np.int64) * ns
cap = n_chunks * cd.compress_bound(ns, nc, 2, 1)
h_comp = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
for hb in [int(a) << 20 for a in sys.argv[2:]] or [256 << 20, 512 << 20, 1 << 30, 2 << 30]:
    cd.set_param('host_batch_bytes', hb); cd.set_param('batch_bytes', max(hb, 2 << 30))
    for rep in range(2):
        t = time.perf_counter(); offs = cd.compress_ptr(h_raw.data_ptr(), 0, rows, nc, 2, 1, h_comp.data_ptr(), 0, cap); dt = time.perf_counter() - t
    print('host_batch %5d MB: e2e %.1f ms  %.2f GB/s' % (hb >> 20, dt * 1e3, n_chunks * cb / dt / 1e9), flush=True)
