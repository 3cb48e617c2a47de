#!/usr/bin/env python
"""Development probe: end-to-end (pinned host -> pinned host) decode of reference-written (plain zlib) chunks versus
the sub-batch size of the block-parallel path.  usage: ref_e2e_probe.py [n_chunks]"""
import sys, time, zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, synth
from oracle import codec as ora
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 600
ns, nc = 30000, 385
cd = _native.default_codec(0)
base = [synth.ap_chunk(ns, nc, seed=100 + i) for i in range(8)]
with ThreadPoolExecutor(8) as ex:
    zs = list(ex.map(lambda x: zlib.compress(ora.transform_chunk(x, True, False)), base))
offs = np.zeros(n_chunks + 1, dtype=np.int64)
offs[1:] = np.cumsum([len(zs[i % 8]) for i in range(n_chunks)])
h_comp = torch.empty(int(offs[-1]), dtype=torch.uint8, pin_memory=True)
for i in range(n_chunks): h_comp.numpy()[offs[i]:offs[i + 1]] = np.frombuffer(zs[i % 8], dtype=np.uint8)
raw = n_chunks * ns * nc * 2
h_out = torch.empty(raw, dtype=torch.uint8, pin_memory=True)
rows = np.arange(n_chunks + 1, dtype=np.int64) * ns
for pb in [int(a) << 20 for a in sys.argv[2:]] or [16 << 30, 7 << 30, 4 << 30, 2 << 30]:
    cd.set_param('par_batch_bytes', pb)
    for rep in range(2):
        t = time.perf_counter()
        cd.decompress_ptr(h_comp.data_ptr(), 0, offs, rows, nc, 2, 1, h_out.data_ptr(), 0)
        dt = time.perf_counter() - t
    print('par_batch %6d MB: e2e %.1f ms  %.2f GB/s  stages %s' % (pb >> 20, dt * 1e3, raw / dt / 1e9, ['%.0f' % v for v in cd.timings()]), flush=True)
ok = all(bytes(h_out.numpy()[i * ns * nc * 2:(i + 1) * ns * nc * 2]) == base[i % 8].tobytes() for i in (0, 7, n_chunks // 2, n_chunks - 1))
print('exact' if ok else 'MISMATCH')
