#!/usr/bin/env python
"""Development probe for BASELINE configs[3]: LFP band (385 ch, 2.5 kHz, 1 s chunks of 1.9 MB, do_spatial_diff=True) —
many small chunks.  Device-resident compress / decompress (own and reference-written streams) with exactness checks.
usage: lfp_probe.py [n_chunks]"""
import sys, time, zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, synth
from oracle import codec as ora
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
ns, nc = 2500, 385
cd = _native.default_codec(0)
fl = _native.flags_of(True, True, 'F')
base = [np.ascontiguousarray(synth.lfp_chunk(ns, nc, seed=50 + i)) for i in range(8)]
with ThreadPoolExecutor(8) as ex:
    zs = list(ex.map(lambda x: ora.encode_chunk(x, True, True, 'F'), base))
cb = ns * nc * 2
raw = np.concatenate([base[i % 8] for i in range(n_chunks)])
rows = np.arange(n_chunks + 1, dtype=np.int64) * ns
d_raw = torch.from_numpy(raw.view(np.uint8).reshape(-1)).cuda()
cap = n_chunks * cd.compress_bound(ns, nc, 2, fl)
d_comp = torch.empty(cap, dtype=torch.uint8, device='cuda')
d_out = torch.empty(n_chunks * cb, dtype=torch.uint8, device='cuda')
ref_offs = np.zeros(n_chunks + 1, dtype=np.int64); ref_offs[1:] = np.cumsum([len(zs[i % 8]) for i in range(n_chunks)])
d_ref = torch.from_numpy(np.frombuffer(b''.join(zs[i % 8] for i in range(n_chunks)), np.uint8).copy()).cuda()
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best, r
t, offs = timed(lambda: cd.compress_ptr(d_raw.data_ptr(), 1, rows, nc, 2, fl, d_comp.data_ptr(), 1, cap))
ref_total = int(ref_offs[-1])
print('LFP %d chunks (%.2f GB): compress %.1f ms = %.1f GB/s, size/zlib %.4f' % (n_chunks, n_chunks * cb / 1e9, t * 1e3, n_chunks * cb / t / 1e9, int(offs[-1]) / ref_total), [round(v, 1) for v in cd.timings()])
t, st = timed(lambda: cd.decompress_ptr(d_comp.data_ptr(), 1, offs, rows, nc, 2, fl, d_out.data_ptr(), 1))
print('  decode own %.1f ms = %.1f GB/s exact %s' % (t * 1e3, n_chunks * cb / t / 1e9, bool(torch.equal(d_out, d_raw)) and not st.any()), [round(v, 1) for v in cd.timings()])
d_out.zero_()
t, st = timed(lambda: cd.decompress_ptr(d_ref.data_ptr(), 1, ref_offs, rows, nc, 2, fl, d_out.data_ptr(), 1))
print('  decode ref %.1f ms = %.1f GB/s exact %s' % (t * 1e3, n_chunks * cb / t / 1e9, bool(torch.equal(d_out, d_raw)) and not st.any()), [round(v, 1) for v in cd.timings()], 'chained', cd.get_param('par_chained'), 'resumed', cd.get_param('par_resumed'))
