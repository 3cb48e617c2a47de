#!/usr/bin/env python
"""Development: device-resident decode of reference-written (plain zlib) chunks, N chunks tiled from a few distinct."""
import sys, time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, synth
from oracle import codec as ora
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nd = min(n_chunks, int(sys.argv[2]) if len(sys.argv) > 2 else 4)
ns, nc = 30000, 385
cd = _native.default_codec(0)
for kv in sys.argv[3:]:
    k, v = kv.split('='); cd.set_param(k, int(v))
base = [synth.ap_chunk(ns, nc, seed=100 + i) for i in range(nd)]
with ThreadPoolExecutor(8) as ex: ref = list(ex.map(ora.encode_chunk, base))
rcomp = np.frombuffer(b''.join(ref[i % nd] for i in range(n_chunks)), dtype=np.uint8)
roffs = np.concatenate(([0], np.cumsum([len(ref[i % nd]) for i in range(n_chunks)]))).astype(np.int64)
rows = np.arange(n_chunks + 1, dtype=np.int64) * ns
lib = cd.lib
raw_bytes = n_chunks * ns * nc * 2
d_c = lib.mtsb_device_alloc(cd.ctx, rcomp.size + 64); d_o = lib.mtsb_device_alloc(cd.ctx, raw_bytes)
lib.mtsb_memcpy(cd.ctx, d_c, rcomp.ctypes.data, rcomp.size, 1)
for it in range(3):
    t = time.time(); st = cd.decompress_ptr(d_c, 1, roffs, rows, nc, 2, 1, d_o, 1); dt = time.time() - t
    print('decode %d ref chunks: %.1f ms  %.2f GB/s  stages %s  par: surv %d cand %d chained %d resumed %d' % (
        n_chunks, dt * 1e3, raw_bytes / dt / 1e9, ['%.1f' % v for v in cd.timings()], cd.get_param('par_survivors'),
        cd.get_param('par_candidates'), cd.get_param('par_chained'), cd.get_param('par_resumed')), flush=True)
assert not st.any()
back = np.empty((ns, nc), np.int16)
lib.mtsb_memcpy(cd.ctx, back.ctypes.data, d_o + (n_chunks - 1) * ns * nc * 2, back.nbytes, 2)
assert np.array_equal(back, base[(n_chunks - 1) % nd]); print('last chunk exact')
