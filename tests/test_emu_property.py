"""CPU: property tests of the kernels' logic under the host emulation (csrc/emu — development aid, see its banner; the
B200 runs of the same properties are tests/test_gpu_parity*.py).  Random shapes, dtypes, chunkings and difference
settings; for every draw
  * zlib (the reference Reader's decoder) accepts each chunk the emulated encoder writes and returns exactly the bytes
    the reference hands to zlib.compress (oracle.transform_chunk, mtscomp.py:381-394);
  * the emulated decoder returns the input, both from its own streams and from reference-written ones (oracle.encode_chunk
    = NumPy diff + zlib.compress, mtscomp.py:375-397);
  * offsets are the reference's chunk_offsets of the batch (monotone, last = total size)."""
import os
import zlib

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import codec as ora


@pytest.fixture(scope='module')
def emu():
    from mtscomp_b200 import _native, build
    return _native.Codec(0, lib=_native.load_library(build.build_emulation()))


DTYPES = ['int8', 'uint8', 'int16', 'uint16', 'int32', 'uint32', 'int64']


@st.composite
def recordings(draw):
    dtype = np.dtype(draw(st.sampled_from(DTYPES)))
    nc = draw(st.one_of(st.integers(1, 40), st.sampled_from([64, 385, 450, 500])))
    n_chunks = draw(st.integers(1, 4))
    max_ns = max(1, min(700, 60000 // (nc * dtype.itemsize)))
    lens = [draw(st.integers(1, max_ns)) for _ in range(n_chunks)]
    seed = draw(st.integers(0, 2 ** 31))
    kind = draw(st.sampled_from(['walk', 'noise', 'const', 'runs', 'extremes']))
    rng = np.random.default_rng(seed)
    ns = sum(lens)
    info = np.iinfo(dtype)
    if kind == 'walk':
        x = np.cumsum(rng.integers(-3, 4, (ns, nc)), axis=0)
    elif kind == 'noise':
        x = rng.integers(info.min, int(info.max) + 1, (ns, nc), dtype=np.int64 if dtype.itemsize < 8 else dtype)
    elif kind == 'const':
        x = np.full((ns, nc), draw(st.integers(-5, 5)))
    elif kind == 'runs':
        x = np.repeat(rng.integers(-50, 50, (ns // 16 + 1, nc)), 16, axis=0)[:ns]
    else:
        x = rng.choice(np.array([info.min, info.max, 0, 1], dtype=np.int64 if dtype.itemsize < 8 else dtype), (ns, nc))
    with np.errstate(over='ignore'):
        x = np.ascontiguousarray(np.asarray(x).astype(dtype))
    td, sd = draw(st.sampled_from([(True, False), (True, True), (False, True), (False, False)]))
    order = draw(st.sampled_from(['F', 'F', 'C']))
    return x, np.concatenate(([0], np.cumsum(lens))), td, sd, order


@settings(max_examples=int(os.environ.get('MTS_PROPERTY_EXAMPLES', 30)), deadline=None,
          suppress_health_check=list(HealthCheck), derandomize='MTS_PROPERTY_EXAMPLES' not in os.environ)
@given(recordings())
def test_emulated_codec_properties(emu, rec):
    from mtscomp_b200 import _native
    x, rows, td, sd, order = rec
    fl = _native.flags_of(td, sd, order)
    kw = dict(do_time_diff=td, do_spatial_diff=sd, chunk_order=order)
    n = len(rows) - 1
    comp, offs = emu.compress(x, rows, fl)
    assert offs[0] == 0 and (np.diff(offs) > 0).all() and offs[-1] == len(comp)
    for i in range(n):
        assert zlib.decompress(bytes(comp[offs[i]:offs[i + 1]])) == ora.transform_chunk(x[rows[i]:rows[i + 1]], **kw), i
    out, stt = emu.decompress(comp, offs, rows, x.shape[1], x.dtype, fl)
    assert not stt.any() and out.dtype == x.dtype and np.array_equal(out, x)
    ref = [ora.encode_chunk(x[rows[i]:rows[i + 1]], **kw) for i in range(n)]
    roffs = np.concatenate(([0], np.cumsum([len(c) for c in ref])))
    out, stt = emu.decompress(b''.join(ref), roffs, rows, x.shape[1], x.dtype, fl)
    assert not stt.any() and np.array_equal(out, x)
