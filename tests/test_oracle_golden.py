"""CPU: pin the oracle (oracle/codec.py, oracle/c) against golden vectors written by the unmodified reference
(tools/make_golden.py imports /root/reference/mtscomp.py; the fixtures travel, the reference does not)."""
import hashlib
import json
import zlib

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import codec as ora

MANIFEST = json.loads((GOLDEN / 'manifest.json').read_text())
CASES = sorted(MANIFEST['cases'])


def load_case(name):
    m = MANIFEST['cases'][name]
    ch = json.loads((GOLDEN / (name + '.ch')).read_text())
    raw = np.fromfile(GOLDEN / (name + '.bin'), dtype=m['dtype']).reshape(m['shape'])
    cbin = (GOLDEN / (name + '.cbin')).read_bytes()
    return m, ch, raw, cbin


def flags_of(ch):
    return dict(do_time_diff=ch['do_time_diff'], do_spatial_diff=ch['do_spatial_diff'], chunk_order=ch['chunk_order'])


@pytest.mark.parametrize('name', CASES)
def test_oracle_decodes_reference_files(name):
    m, ch, raw, cbin = load_case(name)
    assert hashlib.sha1(cbin).hexdigest() == m['sha1_cbin'] == ch['sha1_compressed']
    out = ora.decode_array(cbin, ch['chunk_bounds'], ch['chunk_offsets'], ch['n_channels'], ch['dtype'], **flags_of(ch))
    assert out.dtype == raw.dtype and np.array_equal(out, raw)
    assert hashlib.sha1(out.tobytes()).hexdigest() == m['sha1_decoded'] == ch['sha1_uncompressed']


@pytest.mark.parametrize('name', CASES)
def test_oracle_transform_matches_reference_bytes(name):
    m, ch, raw, _ = load_case(name)
    b = ch['chunk_bounds']
    tr = ora.transform_chunk(raw[b[0]:b[1]], **flags_of(ch))
    assert tr == (GOLDEN / (name + '.tr')).read_bytes()
    assert hashlib.sha1(tr).hexdigest() == m['sha1_tr0']
    back = ora.untransform_bytes(tr, b[1] - b[0], ch['n_channels'], ch['dtype'], **flags_of(ch))
    assert np.array_equal(back, raw[b[0]:b[1]])


@pytest.mark.parametrize('name', CASES)
def test_oracle_encoder_reproduces_reference_cbin(name):
    m, ch, raw, cbin = load_case(name)
    got, bounds, offsets = ora.encode_array(raw, ch['sample_rate'], m['kwargs'].get('chunk_duration', 1.0),
                                            n_threads=2, **flags_of(ch))
    assert bounds == ch['chunk_bounds']
    # the streams are a function of the zlib build; identical bytes are only promised for the recorded version
    if zlib.ZLIB_RUNTIME_VERSION == MANIFEST['zlib_version']:
        assert got == cbin and offsets == ch['chunk_offsets']
    out = ora.decode_array(got, bounds, offsets, ch['n_channels'], ch['dtype'], **flags_of(ch))
    assert np.array_equal(out, raw)


def test_wraparound_semantics():
    x = np.array([[32767, -32768], [-32768, 32767], [0, -1]], dtype=np.int16)
    d = ora.diff_along_axis(x, 0)
    assert d.tolist() == [[32767, -32768], [1, -1], [-32768, -32768]]
    assert np.array_equal(ora.cumsum_along_axis(d, 0), x)
    assert np.array_equal(ora.cumsum_along_axis(ora.diff_along_axis(x, 1), 1), x)


def test_adler_combine_matches_zlib():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, 70001, dtype=np.uint8).tobytes()
    b = rng.integers(0, 256, 123457, dtype=np.uint8).tobytes()
    assert ora.adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)


def test_corrupt_adler_is_rejected():
    _, ch, raw, cbin = load_case('order_c')
    o = ch['chunk_offsets']
    bad = bytearray(cbin[o[0]:o[1]])
    bad[-1] ^= 0x55
    with pytest.raises(zlib.error):
        ora.decode_chunk(bytes(bad), ch['chunk_bounds'][1], ch['n_channels'], ch['dtype'], **flags_of(ch))


# ---- the dependency-free C restatement (oracle/c/mtsoracle.c) against the same reference-written fixtures
@pytest.mark.parametrize('name', CASES)
def test_c_oracle_on_reference_files(name):
    from oracle import cport
    m, ch, raw, cbin = load_case(name)
    b, o = ch['chunk_bounds'], ch['chunk_offsets']
    kw = dict(td=ch['do_time_diff'], sd=ch['do_spatial_diff'], order=ch['chunk_order'])
    assert cport.transform(raw[b[0]:b[1]], **kw) == (GOLDEN / (name + '.tr')).read_bytes()
    for i in range(len(b) - 1):
        n = (b[i + 1] - b[i]) * ch['n_channels'] * raw.dtype.itemsize
        tr = cport.inflate(cbin[o[i]:o[i + 1]], n)
        assert tr == zlib.decompress(cbin[o[i]:o[i + 1]])
        assert cport.adler32(tr) == zlib.adler32(tr)
        assert np.array_equal(cport.untransform(tr, b[i + 1] - b[i], ch['n_channels'], raw.dtype, **kw), raw[b[i]:b[i + 1]])


def test_c_oracle_rejects_bad_adler_and_accepts_trailing_bytes():
    from oracle import cport
    data = bytes(range(256)) * 40
    z = zlib.compress(data)
    assert cport.inflate(z + b'\x01\x02\x03', len(data)) == data
    bad = bytearray(z)
    bad[-2] ^= 1
    with pytest.raises(ValueError):
        cport.inflate(bytes(bad), len(data))
    for level, strategy in [(0, 0), (1, 0), (9, 0), (6, zlib.Z_FIXED)]:
        c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        assert cport.inflate(c.compress(data) + c.flush(), len(data)) == data


# ---- floating point: the parity target is what the reference READER returns (<name>.dec), which differs from the
#      input in the last bits when the differences are not exact (float32_wild)
FLOAT_CASES = sorted(MANIFEST.get('float_cases', {}))


def load_float_case(name):
    m = MANIFEST['float_cases'][name]
    ch = json.loads((GOLDEN / (name + '.ch')).read_text())
    raw = np.fromfile(GOLDEN / (name + '.bin'), dtype=m['dtype']).reshape(m['shape'])
    dec = np.fromfile(GOLDEN / (name + '.dec'), dtype=m['dtype']).reshape(m['shape'])
    return m, ch, raw, dec, (GOLDEN / (name + '.cbin')).read_bytes()


@pytest.mark.parametrize('name', FLOAT_CASES)
def test_oracle_float_matches_reference_reader(name):
    m, ch, raw, dec, cbin = load_float_case(name)
    assert hashlib.sha1(cbin).hexdigest() == m['sha1_cbin'] == ch['sha1_compressed']
    out = ora.decode_array(cbin, ch['chunk_bounds'], ch['chunk_offsets'], ch['n_channels'], ch['dtype'], **flags_of(ch))
    assert out.dtype == dec.dtype and out.tobytes() == dec.tobytes()
    assert np.array_equal(out, raw) == m['exact_round_trip']
    b = ch['chunk_bounds']
    assert ora.transform_chunk(raw[b[0]:b[1]], **flags_of(ch)) == (GOLDEN / (name + '.tr')).read_bytes()
    if zlib.ZLIB_RUNTIME_VERSION == MANIFEST['zlib_version']:
        got, bounds, offsets = ora.encode_array(raw, ch['sample_rate'], 1.0, n_threads=1, **flags_of(ch))
        assert got == cbin and offsets == ch['chunk_offsets']
