#!/usr/bin/env python
"""Quick device-resident timing of the codec stages (development; the contract benchmark is bench.py)."""
import ctypes as C
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, synth  # noqa: E402
from oracle import codec as ora  # noqa: E402


def main():
    n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    n_distinct = min(n_chunks, int(sys.argv[2]) if len(sys.argv) > 2 else 4)
    ns, nc = 30000, 385
    cd = _native.default_codec(0)
    for kv in sys.argv[3:]:
        k, v = kv.split('=')
        cd.set_param(k, int(v))
    t = time.time()
    base = [synth.ap_chunk(ns, nc, seed=100 + i) for i in range(n_distinct)]
    x = np.ascontiguousarray(np.concatenate([base[i % n_distinct] for i in range(n_chunks)], axis=0))
    assert x.flags.c_contiguous
    rows = np.arange(n_chunks + 1, dtype=np.int64) * ns
    print('gen %.1fs' % (time.time() - t), x.shape, flush=True)
    fl = _native.TIME_DIFF
    lib = cd.lib
    raw_bytes = x.nbytes
    d_raw = lib.mtsb_device_alloc(cd.ctx, raw_bytes)
    cap = sum(cd.compress_bound(ns, nc, 2, fl) for _ in range(n_chunks))
    d_comp = lib.mtsb_device_alloc(cd.ctx, cap)
    d_out = lib.mtsb_device_alloc(cd.ctx, raw_bytes)
    lib.mtsb_memcpy(cd.ctx, d_raw, x.ctypes.data, raw_bytes, 1)
    for it in range(3):
        t = time.time()
        offs = cd.compress_ptr(d_raw, 1, rows, nc, 2, fl, d_comp, 1, cap)
        dt = time.time() - t
        tm = cd.timings()
        print('compress  wall %.1f ms  %.2f GB/s  stages(h2d,transform,adler,lz77,huff+scan,encode,d2h,total)=%s launches=%d' % (
            dt * 1e3, raw_bytes / dt / 1e9, ['%.2f' % v for v in tm], cd.launches()), flush=True)
    if hasattr(lib, 'mtsb_debug_lz_profile'):
        buf = (C.c_ulonglong * 16)()
        lib.mtsb_debug_lz_profile(buf)
        v = list(buf); st = max(v[0], 1)
        print('lz profile (cycles/step, thread 0 of CTA 0): steps %d  lookup+compare+stage1 %.0f  insert+emit %.0f  settle %.0f  whole step %.0f' % (v[0], v[1] / st, v[2] / st, v[3] / st, v[5] / st), flush=True)
    csize = int(offs[-1])
    comp = np.empty(csize, dtype=np.uint8)
    lib.mtsb_memcpy(cd.ctx, comp.ctypes.data, d_comp, csize, 2)
    # CPU zlib of the distinct chunks (threads), for ratio and reference-written decode
    t = time.time()
    with ThreadPoolExecutor(8) as ex:
        ref = list(ex.map(lambda b: ora.encode_chunk(b), base))
    tz = time.time() - t
    ref_total = sum(len(ref[i % n_distinct]) for i in range(n_chunks))
    print('ratio gpu %.4f  zlib %.4f  size/zlib %.4f   (zlib %d chunks in %.1fs on 8 threads)' % (
        csize / raw_bytes, ref_total / raw_bytes, csize / ref_total, n_distinct, tz), flush=True)
    # verify a few chunks with zlib
    for i in (0, n_chunks - 1):
        got = zlib.decompress(bytes(comp[offs[i]:offs[i + 1]]))
        assert got == ora.transform_chunk(x[rows[i]:rows[i + 1]]), i
    print('zlib accepts GPU streams', flush=True)
    for it in range(3):
        t = time.time()
        st = cd.decompress_ptr(d_comp, 1, offs, rows, nc, 2, fl, d_out, 1)
        dt = time.time() - t
        tm = cd.timings()
        print('decompress(own)  wall %.1f ms  %.2f GB/s  stages(h2d,plan,inflate,adler,inverse,d2h,-,total)=%s' % (
            dt * 1e3, raw_bytes / dt / 1e9, ['%.2f' % v for v in tm]), flush=True)
    assert not st.any()
    back = np.empty_like(x)
    lib.mtsb_memcpy(cd.ctx, back.ctypes.data, d_out, raw_bytes, 2)
    assert np.array_equal(back, x)
    print('own round trip exact', flush=True)
    # reference-written streams
    rcomp = np.frombuffer(b''.join(ref[i % n_distinct] for i in range(n_chunks)), dtype=np.uint8)
    roffs = np.concatenate(([0], np.cumsum([len(ref[i % n_distinct]) for i in range(n_chunks)]))).astype(np.int64)
    d_rcomp = lib.mtsb_device_alloc(cd.ctx, rcomp.size + 64)
    lib.mtsb_memcpy(cd.ctx, d_rcomp, rcomp.ctypes.data, rcomp.size, 1)
    for it in range(2):
        t = time.time()
        st = cd.decompress_ptr(d_rcomp, 1, roffs, rows, nc, 2, fl, d_out, 1)
        dt = time.time() - t
        tm = cd.timings()
        print('decompress(ref)  wall %.1f ms  %.2f GB/s  stages=%s' % (dt * 1e3, raw_bytes / dt / 1e9, ['%.2f' % v for v in tm]), flush=True)
    assert not st.any()
    lib.mtsb_memcpy(cd.ctx, back.ctypes.data, d_out, raw_bytes, 2)
    assert np.array_equal(back, x)
    print('reference-written decode exact', flush=True)


if __name__ == '__main__':
    main()
