#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference/mtscomp.py).

Run in the build container only (the reference does not exist on the GPU box):  python tools/make_golden.py
Each case writes <name>.bin (raw input), <name>.cbin / <name>.ch (reference Writer output) and, for the transform
kernels, <name>.tr (the bytes the reference hands to zlib for chunk 0).  manifest.json records shapes, flags, the
zlib version that produced the streams and sha1 of what the reference Reader returns.
"""
import hashlib
import importlib.util
import json
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import synth  # noqa: E402

spec = importlib.util.spec_from_file_location('mtscomp_reference', '/root/reference/mtscomp.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
ref.CONFIG_PATH = Path('/nonexistent/.mtscomp')

OUT = ROOT / 'tests' / 'golden'
OUT.mkdir(parents=True, exist_ok=True)


def cases():
    rng = np.random.default_rng(7)
    # (name, array, sample_rate, kwargs)
    yield 'ap_small', synth.ap_chunk(ns=700, nc=385, sample_rate=30000., seed=11), 300., {}
    yield 'lfp_spatial', synth.lfp_chunk(ns=900, nc=97, sample_rate=2500., seed=12), 400., dict(do_spatial_diff=True)
    yield 'order_c', synth.ap_chunk(ns=500, nc=64, sample_rate=30000., seed=13), 200., dict(chunk_order='C')
    yield 'fullrange', rng.integers(-32768, 32768, (300, 33), dtype=np.int64).astype(np.int16), 128., {}
    edge = np.array([[32767, -32768, 0], [-32768, 32767, -1], [32767, 32767, 1], [0, -1, -32768]], dtype=np.int16)
    yield 'wrap_edges', np.tile(edge, (25, 5)), 30., {}
    yield 'tiny_1x1', np.array([[-1234]], dtype=np.int16), 1., {}
    yield 'tiny_chunks', synth.ap_chunk(ns=100, nc=19, sample_rate=1234., seed=14, rms=300.), 1234., dict(chunk_duration=.01)
    z = np.zeros((2000, 16), dtype=np.int16)
    z[:, 3] = 17
    z[1000:, 5] = -3
    yield 'zeros_runs', z, 1000., {}
    yield 'spatial_only', synth.lfp_chunk(ns=400, nc=40, sample_rate=2500., seed=15), 150., dict(do_time_diff=False, do_spatial_diff=True)
    yield 'nodiff', synth.ap_chunk(ns=300, nc=24, sample_rate=30000., seed=16), 100., dict(do_time_diff=False)
    yield 'one_channel', synth.ap_chunk(ns=5000, nc=2, sample_rate=30000., seed=17)[:, :1].copy(), 2000., {}


def main():
    manifest = {'zlib_version': zlib.ZLIB_RUNTIME_VERSION, 'numpy_version': np.__version__,
                'reference_version': ref.__version__, 'cases': {}}
    for name, arr, sr, kw in cases():
        arr = np.ascontiguousarray(arr)
        raw = OUT / (name + '.bin')
        arr.tofile(raw)
        cbin, ch = OUT / (name + '.cbin'), OUT / (name + '.ch')
        ref.compress(raw, cbin, ch, sample_rate=sr, n_channels=arr.shape[1], dtype=arr.dtype,
                     n_threads=1, check_after_compress=True, quiet=True, **kw)
        r = ref.decompress(cbin, ch)
        dec = r[:]
        assert np.array_equal(dec, arr)
        # transform bytes for chunk 0, exactly as mtscomp.py:381-394 builds them
        b = r.chunk_bounds
        c0 = arr[b[0]:b[1]]
        d = ref.diff_along_axis(c0, axis=0 if r.cmeta.do_time_diff else None)
        d = ref.diff_along_axis(d, axis=1 if r.cmeta.do_spatial_diff else None)
        tr = d.tobytes(order=r.cmeta.chunk_order)
        (OUT / (name + '.tr')).write_bytes(tr)
        manifest['cases'][name] = dict(
            shape=list(arr.shape), dtype=str(arr.dtype), sample_rate=sr, kwargs=kw,
            n_chunks=r.n_chunks, cbin_bytes=cbin.stat().st_size,
            sha1_decoded=hashlib.sha1(dec.tobytes()).hexdigest(),
            sha1_cbin=hashlib.sha1(cbin.read_bytes()).hexdigest(),
            sha1_tr0=hashlib.sha1(tr).hexdigest())
        r.close()
        print(name, arr.shape, kw, 'chunks', manifest['cases'][name]['n_chunks'], 'cbin', cbin.stat().st_size)
    (OUT / 'manifest.json').write_text(json.dumps(manifest, indent=1, sort_keys=True))


def float_cases():
    rng = np.random.default_rng(21)
    t = np.arange(1500) / 1000.
    base = 300. * np.sin(2 * np.pi * 7 * t)[:, None] + rng.normal(0, 20., (1500, 9)).cumsum(axis=0) + 5000.
    yield 'float32_time', base.astype(np.float32), 500., {}
    yield 'float64_spatial_c', base.astype(np.float64), 500., dict(do_spatial_diff=True, chunk_order='C')
    # wide dynamic range: the differences are NOT exact, the reference Reader returns values that differ from the
    # input in the last bits (and its own post-compression check would refuse the file: it is switched off)
    wild = rng.normal(0, 1, (1200, 7)) * 10. ** rng.integers(-3, 5, (1200, 7))
    wild[10, 2] = -0.0
    yield 'float32_wild', wild.astype(np.float32), 400., dict(do_spatial_diff=True, check_after_compress=False)


def main_float():
    """Floating point fixtures (added later; the integer fixtures above are left untouched).  The reference's round trip
    is not exact for floats (cumsum of differences), so the parity target is what the reference READER returns: it is
    stored as <name>.dec next to the reference-written .cbin / .ch."""
    mpath = OUT / 'manifest.json'
    manifest = json.loads(mpath.read_text())
    manifest['float_cases'] = {}
    for name, arr, sr, kw in float_cases():
        arr = np.ascontiguousarray(arr)
        raw = OUT / (name + '.bin')
        arr.tofile(raw)
        cbin, ch = OUT / (name + '.cbin'), OUT / (name + '.ch')
        ref.compress(raw, cbin, ch, sample_rate=sr, n_channels=arr.shape[1], dtype=arr.dtype,
                     n_threads=1, quiet=True, **dict(dict(check_after_compress=True), **kw))
        r = ref.decompress(cbin, ch)
        dec = np.ascontiguousarray(r[:])
        assert dec.dtype == arr.dtype and np.allclose(dec, arr, rtol=1e-3, atol=1.)
        (OUT / (name + '.dec')).write_bytes(dec.tobytes())
        b = r.chunk_bounds
        c0 = arr[b[0]:b[1]]
        d = ref.diff_along_axis(c0, axis=0 if r.cmeta.do_time_diff else None)
        d = ref.diff_along_axis(d, axis=1 if r.cmeta.do_spatial_diff else None)
        tr = d.tobytes(order=r.cmeta.chunk_order)
        (OUT / (name + '.tr')).write_bytes(tr)
        manifest['float_cases'][name] = dict(
            shape=list(arr.shape), dtype=str(arr.dtype), sample_rate=sr,
            kwargs={k: v for k, v in kw.items() if k != 'check_after_compress'}, n_chunks=r.n_chunks,
            cbin_bytes=cbin.stat().st_size, exact_round_trip=bool(np.array_equal(dec, arr)),
            sha1_decoded=hashlib.sha1(dec.tobytes()).hexdigest(),
            sha1_cbin=hashlib.sha1(cbin.read_bytes()).hexdigest(), sha1_tr0=hashlib.sha1(tr).hexdigest())
        r.close()
        print(name, arr.shape, kw, 'chunks', r.n_chunks, 'cbin', cbin.stat().st_size, 'exact', manifest['float_cases'][name]['exact_round_trip'])
    mpath.write_text(json.dumps(manifest, indent=1, sort_keys=True))


if __name__ == '__main__':
    import sys
    if 'float' in sys.argv[1:]:
        main_float()
    else:
        main()
