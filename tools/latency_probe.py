#!/usr/bin/env python
"""Development probe: where the time of one cold Reader slice goes (pread / decode on the device / D2H / Python)."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tools'))
import mtscomp_b200 as M
from mtscomp_b200 import _native, synth
import bench_legs

ns, nc, n_chunks = 30000, 384, 8
d = bench_legs.scratch_dir('probe')
M.CONFIG_PATH = d / '.mtscomp'
base = [synth.ap_chunk(ns, nc, seed=300 + i) for i in range(4)]
arr = np.concatenate([base[i % 4] for i in range(n_chunks)])
arr.tofile(d / 'np2.bin')
bench_legs.write_reference_style(arr, 30000., d / 'ref.cbin', d / 'ref.ch', 16)
M.compress(d / 'np2.bin', d / 'gpu.cbin', d / 'gpu.ch', sample_rate=30000., n_channels=nc, dtype=np.int16, quiet=True, check_after_compress=False)
cd = _native.default_codec()
for name in ('ref', 'gpu'):
    r = M.decompress(d / (name + '.cbin'), d / (name + '.ch'), cache_size=1)
    r[0:300]
    for rep in range(3):
        for idx in (3, 5, 6):
            t0 = time.perf_counter()
            start, length = r._span(idx)
            buf = cd.host_buffer('r_comp', length + 64)
            r._pread_into(buf.array[:length], start)
            t1 = time.perf_counter()
            blk = M.core._DeviceBlock(cd, ns * nc * 2 + 256)
            t2 = time.perf_counter()
            rows = np.array([0, ns], dtype=np.int64)
            st = cd.decompress_ptr(buf.ptr, 0, np.array([0, length], dtype=np.int64), rows, nc, 2, r._flags(), blk.ptr, 1)
            t3 = time.perf_counter()
            tm = cd.timings()
            out = np.empty((300, nc), np.int16)
            cd.memcpy(out.ctypes.data, blk.ptr + 1000 * nc * 2, out.nbytes, 2)
            t4 = time.perf_counter()
            del blk
            t5 = time.perf_counter()
            a = r[idx * ns + 1000: idx * ns + 1300]
            t6 = time.perf_counter()
            assert np.array_equal(out, arr[idx * ns + 1000: idx * ns + 1300]) and np.array_equal(a, out)
            print('%s chunk %d: pread %.2f  alloc %.2f  decode call %.2f (stages %s)  d2h %.2f  free %.2f | r[...] %.2f ms' % (
                name, idx, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, ['%.2f' % v for v in tm], (t4 - t3) * 1e3, (t5 - t4) * 1e3, (t6 - t5) * 1e3), flush=True)
    r.close()
import shutil; shutil.rmtree(d, ignore_errors=True)
