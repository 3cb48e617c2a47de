"""Drop-in API behaviours, mirroring the reference's own tests (reference tests.py) on the integer codec path.
Each function is run twice: on the B200 (tests/test_gpu_api.py) and through the host emulation of the kernels
(tests/test_emu_api.py, CPU container)."""
import hashlib
import io
import json
import re
from contextlib import redirect_stdout
from itertools import product
from pathlib import Path

import numpy as np
from pytest import raises

import mtscomp_b200 as M
from mtscomp_b200 import (
    Reader, Writer, compress, cumsum_along_axis, decompress, diff_along_axis, load_raw_data, mtschop, mtscomp,
    mtscomp_parser, mtsdecomp, mtsdecomp_parser, mtsdesc, read_config, _args_to_config)

n_channels = 19
sample_rate = 1234.
duration = 5.67
n_samples = len(np.arange(0, duration, 1. / sample_rate))   # 6997, as reference tests.py:39-44


def _arr16(seed=0, ns=n_samples, nc=n_channels):
    rng = np.random.default_rng(seed)
    t = np.arange(ns) / sample_rate
    a = np.sin(10 * t)[:, None] + rng.normal(0, .25, (ns, nc))
    return (a / np.abs(a).max() * 32766).astype(np.int16)


def _use_tmp_config(tmp_path):
    M.CONFIG_PATH = tmp_path / '.mtscomp'


def _round_trip(tmp_path, arr, **kw):
    _use_tmp_config(tmp_path)
    path = tmp_path / 'data.bin'
    arr.tofile(path)
    out, outmeta = tmp_path / 'data.cbin', tmp_path / 'data.ch'
    compress(path, out, outmeta, sample_rate=sample_rate, n_channels=arr.shape[1], dtype=arr.dtype, **kw)
    unc = decompress(out, outmeta)
    assert np.array_equal(unc[:], arr)
    return unc


def case_diff_cumsum(tmp_path):   # reference tests.py:190-205
    arr = _arr16(1)
    for a1, a2 in product((None, 0, 1), (None, 0, 1)):
        d = diff_along_axis(diff_along_axis(arr, a1), a2)
        back = cumsum_along_axis(cumsum_along_axis(d, a2), a1)
        assert back.dtype == arr.dtype and np.array_equal(back, arr)
    assert np.array_equal(diff_along_axis(arr, 0)[1:], np.diff(arr, axis=0))
    assert np.array_equal(diff_along_axis(arr, 1)[:, 1:], np.diff(arr, axis=1))


def case_low_level(tmp_path):   # reference tests.py:212-233
    _use_tmp_config(tmp_path)
    arr = _arr16(2)
    path = tmp_path / 'data.bin'
    arr.tofile(path)
    w = Writer()
    w.open(path, sample_rate=sample_rate, n_channels=arr.shape[1], dtype=arr.dtype)
    assert w.n_chunks == 6 and w.chunk_bounds[-1] == n_samples
    ratio = w.write(None, None)
    w.close()
    out, outmeta = tmp_path / 'data.cbin', tmp_path / 'data.ch'
    assert out.exists() and outmeta.exists()
    assert 0 < ratio < 1 and abs(ratio - out.stat().st_size / arr.nbytes) < 1e-12
    r = Reader()
    r.open(out, outmeta)
    assert r.shape == arr.shape and r.dtype == arr.dtype and r.n_chunks == 6
    assert np.array_equal(r[:], arr)
    r.close()


def case_dtypes(tmp_path):   # reference tests.py:240-243 (integer dtypes)
    for dt in ('uint8', 'uint16', 'int8', 'int16', 'int32'):
        info = np.iinfo(dt)
        a = _arr16(3).astype(np.float64) / 32766
        arr = (a * min(info.max, 30000) * (0.5 if info.min == 0 else 1) + (info.max // 2 if info.min == 0 else 0)).astype(dt)
        unc = _round_trip(tmp_path, arr, quiet=True)
        assert unc.dtype == np.dtype(dt)
        unc.close()


def case_float_dtypes(tmp_path):   # reference tests.py:240-243 (floating point dtypes)
    """float32 / float64: as in the reference the round trip is close, not exact (cumsum of differences), and what the
    Reader returns is exactly what the reference Reader returns: np.cumsum's sequential sums of the stored differences."""
    from oracle import codec as ora
    _use_tmp_config(tmp_path)
    for dt, kw in ((np.float32, {}), (np.float64, dict(do_spatial_diff=True)), (np.float32, dict(chunk_order='C'))):
        # (offset: the reference's post-compression check is np.allclose with rtol 1e-5, which float32 sums near zero fail)
        arr = (_arr16(4).astype(np.float64) / 7.3 + 20000.).astype(dt)
        path = tmp_path / 'f.bin'
        arr.tofile(path)
        out, outmeta = tmp_path / 'f.cbin', tmp_path / 'f.ch'
        compress(path, out, outmeta, sample_rate=sample_rate, n_channels=arr.shape[1], dtype=arr.dtype, quiet=True, **kw)
        unc = decompress(out, outmeta)
        got = unc[:]
        assert got.dtype == np.dtype(dt) and np.allclose(got, arr, atol=0.5 if dt == np.float32 else 1e-8)
        b = unc.chunk_bounds
        flags = dict(do_time_diff=True, do_spatial_diff=kw.get('do_spatial_diff', False), chunk_order=kw.get('chunk_order', 'F'))
        want = np.concatenate([ora.decode_chunk(ora.encode_chunk(arr[b[i]:b[i + 1]], **flags), b[i + 1] - b[i],
                                                arr.shape[1], dt, **flags) for i in range(len(b) - 1)])
        assert got.tobytes() == want.tobytes()
        unc.close()
    with raises(NotImplementedError):
        _round_trip(tmp_path, np.zeros((100, 4), np.float16), quiet=True)


def case_comp_decomp_hashes(tmp_path):   # reference tests.py:381-410
    _use_tmp_config(tmp_path)
    rng = np.random.default_rng(5)
    arr = np.cumsum(rng.integers(-20, 21, (1000, 1000)), axis=0).astype(np.int16)
    path = tmp_path / 'data.bin'
    arr.tofile(path)
    out, outmeta, back = tmp_path / 'data.cbin', tmp_path / 'data.ch', tmp_path / 'back.bin'
    compress(path, out, outmeta, sample_rate=100., n_channels=1000, dtype=np.int16, quiet=True)
    decompress(out, outmeta, back, quiet=True).close()
    assert back.read_bytes() == path.read_bytes()
    meta = json.loads(outmeta.read_text())
    assert meta['sha1_compressed'] == hashlib.sha1(out.read_bytes()).hexdigest()
    assert meta['sha1_uncompressed'] == hashlib.sha1(path.read_bytes()).hexdigest()
    assert list(meta.keys()) == sorted(meta.keys())
    assert set(meta) == {'version', 'algorithm', 'comp_level', 'do_time_diff', 'do_spatial_diff', 'dtype', 'n_channels',
                         'sample_rate', 'chunk_bounds', 'chunk_offsets', 'chunk_order', 'sha1_compressed',
                         'sha1_uncompressed', 'shape'}
    assert meta['chunk_offsets'][-1] == out.stat().st_size and meta['version'] == '1.0'


def case_parameters(tmp_path):   # reference tests.py:499-526
    arr = _arr16(6)
    for cd in (.01, .1, 1, 10):
        _round_trip(tmp_path, arr, chunk_duration=cd, quiet=True).close()
    for td, sd in product((False, True), (False, True)):
        _round_trip(tmp_path, arr, do_time_diff=td, do_spatial_diff=sd, comp_level=3, quiet=True).close()
    for order in 'FC':
        _round_trip(tmp_path, arr, chunk_order=order, quiet=True).close()
    for nt in (1, 2, 4, None):
        _round_trip(tmp_path, arr, n_threads=nt, quiet=True).close()


def case_n_channels(tmp_path):   # reference tests.py:504-512
    for ns, nc in product((0, 1, 100, 10000), (0, 1, 10, 100)):
        arr = _arr16(7, max(ns, 1), max(nc, 1))[:ns, :nc]
        if ns * nc == 0:
            with raises(Exception):
                _round_trip(tmp_path, arr, quiet=True)
        else:
            _round_trip(tmp_path, arr, quiet=True).close()


def case_reader_indexing(tmp_path):   # reference tests.py:246-342
    arr = _arr16(8)
    unc = _round_trip(tmp_path, arr, quiet=True)
    N = n_samples
    items = [slice(a, b, c) for a, b, c in product((None, 0, 1, -1), (None, 0, 1, -1), (None, 2, 3, N // 2, N))]
    X = np.random.default_rng(1).integers(-100, 2 * N, (100, 3))
    items += [slice(int(a), int(b), int(c)) for a, b, c in X]
    items += [(slice(None),), (slice(None), slice(1, -1, 2)), (slice(None), [1, 5, 3]), (slice(None), 1),
              (1, slice(None)), (2, 1), 0, 1, N - 2, N - 1]
    items += np.random.default_rng(2).integers(-N, N, 100).tolist()
    for t1, t2 in product([np.uint64, np.int64, np.int8, int], repeat=2):
        items.append(slice(t1(1), t2(3)))
        items.append(slice(t1(5), t2(9), np.int64(2)))
    for s in items:
        if isinstance(s, slice) and s.step is not None and s.step <= 0:
            continue
        got, want = unc[s], arr[s]
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want)
    # exact chunk selection, as the reference pins it (tests.py:308-342): bounds 0,1234,...,6170,6997
    b = unc.chunk_bounds
    assert b == [0, 1234, 2468, 3702, 4936, 6170, 6997]
    assert unc._chunks_for_interval(0, 1) == (0, 0)
    assert unc._chunks_for_interval(0, 1233) == (0, 0)
    assert unc._chunks_for_interval(0, 1234) == (0, 1)
    assert unc._chunks_for_interval(1233, 1234) == (0, 1)
    assert unc._chunks_for_interval(1234, 1235) == (1, 1)
    assert unc._chunks_for_interval(2467, 2468) == (1, 2)
    assert unc._chunks_for_interval(6996, 6997) == (5, 5)
    assert unc._chunks_for_interval(6997, 6998) == (5, 5)
    unc.close()


def case_check_fail(tmp_path):   # reference tests.py:345-378
    _use_tmp_config(tmp_path)
    arr = _arr16(9)
    path = tmp_path / 'data.bin'
    arr.tofile(path)

    def corrupt(writer):
        with open(path, 'r+b') as f:
            f.seek(5000)
            f.write(b'\x12\x34\x56\x78\x9a\xbc\xde\xf0')
    w = Writer(before_check=corrupt, quiet=True)
    w.open(path, sample_rate=sample_rate, n_channels=n_channels, dtype=np.int16)
    with raises(RuntimeError):
        w.write(tmp_path / 'data.cbin', tmp_path / 'data.ch')
    w.close()


def case_corrupt_chunk_is_ioerror(tmp_path):   # reference mtscomp.py:618-621
    arr = _arr16(10)
    unc = _round_trip(tmp_path, arr, quiet=True)
    unc.close()
    out = tmp_path / 'data.cbin'
    b = bytearray(out.read_bytes())
    meta = json.loads((tmp_path / 'data.ch').read_text())
    b[meta['chunk_offsets'][2] + 100] ^= 0xff
    out.write_bytes(bytes(b))
    r = decompress(out, tmp_path / 'data.ch')
    assert np.array_equal(r[:1234], arr[:1234])
    with raises(IOError, match='Compressed chunk #2 is corrupted'):
        r[2468:2500]
    r.close()


def case_decompress_pool(tmp_path):   # reference tests.py:413-430
    arr = _arr16(11)
    unc = _round_trip(tmp_path, arr, cache_size=2, quiet=True)
    pool = unc.start_thread_pool()
    got = unc.decompress_chunks([0, 2, 5], pool)
    unc.stop_thread_pool()
    b = unc.chunk_bounds
    assert sorted(got) == [0, 2, 5]
    for i, c in got.items():
        assert np.array_equal(c, arr[b[i]:b[i + 1]])
    unc.close()


def case_npy_3d(tmp_path):   # reference tests.py:433-448
    _use_tmp_config(tmp_path)
    arr = np.random.default_rng(12).integers(-3000, 3000, (100, 20, 10)).astype(np.int16)
    path = tmp_path / 'data.npy'
    np.save(path, arr)
    compress(path, sample_rate=20., quiet=True)
    out, outmeta = tmp_path / 'data.cnpy', tmp_path / 'data.ch'
    assert out.exists() and outmeta.exists()
    r = decompress(out, outmeta)
    assert json.loads(outmeta.read_text())['shape'] == [100, 20, 10]
    assert np.array_equal(r[:].reshape(arr.shape), arr)
    r.close()


def case_chop(tmp_path):   # reference tests.py:451-492
    _use_tmp_config(tmp_path)
    arr = _arr16(13)
    path = tmp_path / 'data.bin'
    arr.tofile(path)
    out, outmeta = tmp_path / 'data.cbin', tmp_path / 'data.ch'
    compress(path, out, outmeta, sample_rate=sample_rate, n_channels=n_channels, dtype=np.int16, quiet=True)
    r = decompress(out, outmeta)
    chopped = tmp_path / 'chop' / 'data.cbin'
    r.chop(2, chopped)
    r.close()
    rc = decompress(chopped)
    assert rc.n_chunks == 2 and rc.cmeta.chopped is True and rc.cmeta.sha1_compressed is None
    assert np.array_equal(rc[:], arr[:2468])
    rc.close()
    # compressing the first two chunks directly gives the same bytes (deterministic, chunk-independent encoder)
    path2 = tmp_path / 'two.bin'
    arr[:2468].tofile(path2)
    compress(path2, tmp_path / 'two.cbin', tmp_path / 'two.ch', sample_rate=sample_rate, n_channels=n_channels,
             dtype=np.int16, quiet=True)
    assert hashlib.sha1((tmp_path / 'two.cbin').read_bytes()).hexdigest() == hashlib.sha1(chopped.read_bytes()).hexdigest()


def case_config_and_cli(tmp_path):   # reference tests.py:152-158, 533-712
    _use_tmp_config(tmp_path)
    cfg = read_config()
    assert cfg.algorithm == 'zlib' and cfg.chunk_duration == 1. and cfg.chunk_order == 'F' and cfg.cache_size == 10
    assert cfg.do_time_diff is True and cfg.do_spatial_diff is False and cfg.comp_level == -1
    p = mtscomp_parser()
    pargs, config = _args_to_config(p, ['somefile', '-d', 'int16', '-s', '30000', '-n', '385', '-c', '2', '-nc', '-p', '3'])
    assert config.dtype == 'int16' and config.sample_rate == 30000 and config.n_channels == 385
    assert config.chunk_duration == 2 and config.check_after_compress is False and config.n_threads == 3
    pargs, config = _args_to_config(mtsdecomp_parser(), ['a.cbin', 'a.ch', '-o', 'x.bin', '-f'], compress=False)
    assert pargs.out == 'x.bin' and pargs.overwrite and config.check_after_decompress is True
    arr = _arr16(14)
    path = tmp_path / 'data.bin'
    arr.tofile(path)
    with raises(ValueError):
        mtscomp([str(path), '-n', str(n_channels), '-d', 'int16'])     # no sample rate
    mtscomp([str(path), '-n', str(n_channels), '-s', str(sample_rate), '-d', 'int16', '--set-default'])
    assert (tmp_path / 'data.cbin').exists() and (tmp_path / '.mtscomp').exists()
    assert read_config().n_channels == n_channels
    (tmp_path / 'data.cbin').unlink()
    (tmp_path / 'data.ch').unlink()
    mtscomp([str(path)])                                                 # defaults come from the config file now
    f = io.StringIO()
    with redirect_stdout(f):
        mtsdesc([str(tmp_path / 'data.cbin')])
    desc = f.getvalue()
    assert re.search(r'n_channels\s+19', desc) and re.search(r'n_chunks\s+6', desc) and 'int16' in desc
    mtsdecomp([str(tmp_path / 'data.cbin'), '-o', str(tmp_path / 'back.bin')])
    assert (tmp_path / 'back.bin').read_bytes() == path.read_bytes()
    mtschop([str(tmp_path / 'data.cbin'), '-n', '3', '-o', str(tmp_path / 'c3.cbin')])
    r = decompress(tmp_path / 'c3.cbin')
    assert np.array_equal(r[:], arr[:3702])
    r.close()


def case_load_raw_data(tmp_path):   # reference tests.py:161-180
    arr = _arr16(15)
    path = tmp_path / 'data.bin'
    arr.tofile(path)
    for mmap in (True, False):
        assert np.array_equal(load_raw_data(path, n_channels=n_channels, dtype=np.int16, mmap=mmap), arr)
    with raises(ValueError):
        load_raw_data(path, n_channels=n_channels + 1, dtype=np.int16)
    (tmp_path / 'empty.bin').write_bytes(b'')
    assert load_raw_data(tmp_path / 'empty.bin', n_channels=3, dtype=np.int16).shape == (0, 3)


def case_write_index_option(tmp_path):   # (no reference counterpart: the in-band index is this package's superset)
    """Default: every chunk is a zlib stream followed by the index (zlib ignores it); write_index=False: the
    reference's exact layout, nothing after the streams.  Both decode here and with CPython zlib."""
    import zlib
    _use_tmp_config(tmp_path)
    arr = _arr16(5)
    arr.tofile(tmp_path / 'data.bin')
    sizes = {}
    for name, kw in (('idx', {}), ('plain', dict(write_index=False))):
        compress(tmp_path / 'data.bin', tmp_path / (name + '.cbin'), tmp_path / (name + '.ch'), sample_rate=sample_rate,
                 n_channels=n_channels, dtype=np.int16, quiet=True, **kw)
        meta = json.loads((tmp_path / (name + '.ch')).read_text())
        blob = (tmp_path / (name + '.cbin')).read_bytes()
        o = meta['chunk_offsets']
        unused = 0
        for i in range(len(o) - 1):
            d = zlib.decompressobj()
            d.decompress(blob[o[i]:o[i + 1]])
            assert d.eof
            unused += len(d.unused_data)
        sizes[name] = (len(blob), unused)
        r = decompress(tmp_path / (name + '.cbin'), tmp_path / (name + '.ch'))
        assert np.array_equal(r[:], arr)
        r.close()
    assert sizes['plain'][1] == 0 and sizes['idx'][1] > 0
    assert sizes['idx'][0] - sizes['idx'][1] == sizes['plain'][0]          # the same streams, plus the index
    assert M._native.default_codec().get_param('write_index') == 1         # the option does not stick to the codec


ALL_CASES = [v for k, v in sorted(globals().items()) if k.startswith('case_')]
