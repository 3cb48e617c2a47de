#!/usr/bin/env python
"""DEVELOPMENT TOOL: run the kernels' logic through the host emulation build (csrc/emu) and check it against the
oracle.  Not a product path; the package never loads the emulation library."""
import sys
import time
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, build, synth  # noqa: E402
from oracle import codec as ora  # noqa: E402


def get_codec():
    lib = _native.load_library(build.build_emulation())
    return _native.Codec(0, lib=lib)


def check_transform(cd, x, td, sd, order):
    fl = _native.flags_of(td, sd, order)
    want = ora.transform_chunk(x, td, sd, order)
    got = cd.delta_transform(x, fl).tobytes()
    assert got == want, ('fwd', x.shape, td, sd, order)
    back, ad = cd.inverse_transform(want, x.shape[0], x.shape[1], x.dtype, fl, want_adler=True)
    assert np.array_equal(back, x), ('inv', x.shape, td, sd, order)
    assert ad == zlib.adler32(want), ('adler', hex(ad), hex(zlib.adler32(want)))


def check_codec(cd, x, rows, td=True, sd=False, order='F', label=''):
    fl = _native.flags_of(td, sd, order)
    t = time.time()
    comp, offs = cd.compress(x, rows, fl)
    t1 = time.time() - t
    ref_total = 0
    for i in range(len(rows) - 1):
        c = bytes(comp[offs[i]:offs[i + 1]])
        want = ora.transform_chunk(x[rows[i]:rows[i + 1]], td, sd, order)
        got = zlib.decompress(c)
        assert got == want, ('zlib roundtrip', label, i)
        ref_total += len(zlib.compress(want))
    t = time.time()
    out, st = cd.decompress(comp, offs, rows, x.shape[1], x.dtype, fl)
    t2 = time.time() - t
    assert not st.any(), st
    assert np.array_equal(out, x), ('own decode', label)
    # reference-written streams through the GPU decoder
    parts = [ora.encode_chunk(x[rows[i]:rows[i + 1]], td, sd, order) for i in range(len(rows) - 1)]
    roffs = np.concatenate(([0], np.cumsum([len(p) for p in parts])))
    out2, st2 = cd.decompress(b''.join(parts), roffs, rows, x.shape[1], x.dtype, fl)
    assert not st2.any(), st2
    assert np.array_equal(out2, x), ('ref decode', label)
    print('%-28s ok  raw %9d  gpu %9d  zlib %9d  size/zlib %.4f  (emu %.1fs enc, %.1fs dec)' % (
        label, x.nbytes, len(comp), ref_total, len(comp) / ref_total, t1, t2))
    return len(comp) / ref_total


def main():
    cd = get_codec()
    rng = np.random.default_rng(0)
    if 'fast' not in sys.argv:
        for shape in [(1, 1), (5, 3), (64, 7), (65, 385), (200, 384), (130, 33)]:
            for dt in (np.int16, np.uint8, np.int32, np.int64):
                info = np.iinfo(dt)
                x = rng.integers(info.min, info.max, shape, dtype=dt, endpoint=True)
                for td in (True, False):
                    for sd in (True, False):
                        for order in 'FC':
                            check_transform(cd, x, td, sd, order)
        print('transforms ok')
    x = synth.ap_chunk(ns=3000, nc=24, seed=5)
    check_codec(cd, x, [0, 3000], label='ap 3000x24 1 chunk')
    x = synth.ap_chunk(ns=700, nc=385, sample_rate=30000., seed=11)
    check_codec(cd, x, [0, 300, 600, 700], label='ap 700x385 3 chunks')
    x = synth.lfp_chunk(ns=900, nc=97, seed=12)
    check_codec(cd, x, [0, 400, 800, 900], sd=True, label='lfp spatial')
    check_codec(cd, x, [0, 900], order='C', label='lfp order C')
    x = rng.integers(-32768, 32767, (300, 33)).astype(np.int16)
    check_codec(cd, x, [0, 128, 256, 300], label='fullrange (stored)')
    z = np.zeros((2000, 16), dtype=np.int16); z[:, 3] = 17; z[1000:, 5] = -3
    check_codec(cd, z, [0, 1000, 2000], label='zeros/runs')
    check_codec(cd, np.array([[-1234]], dtype=np.int16), [0, 1], label='1x1')
    x8 = (synth.ap_chunk(ns=2000, nc=16, seed=3) & 0xff).astype(np.uint8)
    check_codec(cd, x8, [0, 2000], label='uint8')
    x32 = synth.ap_chunk(ns=2000, nc=16, seed=4).astype(np.int32)
    check_codec(cd, x32, [0, 1000, 2000], label='int32')
    if 'big' in sys.argv:
        x = synth.ap_chunk(ns=30000, nc=16, seed=21)
        check_codec(cd, x, [0, 30000], label='ap 30000x16')


if __name__ == '__main__':
    main()
