"""CPU, build container only: the UNMODIFIED reference (`/root/reference/mtscomp.py`) with the reference-side ctypes
binding of INTEGRATION.md §2 installed (`integration/reference_binding.py`): its own compress() / decompress() / Reader
slicing / check run with the codec seam routed through the C ABI — here the host emulation of the kernels, on the GPU
box the same symbols of libmtscomp_b200.so.  The files it writes must open in a second, untouched copy of the reference,
and reference-written files must decode through the binding, byte for byte.  Skipped where /root/reference is absent."""
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path('/root/reference/mtscomp.py')
ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(not REF.exists(), reason='reference tree not present')


def _load_reference(name):
    spec = importlib.util.spec_from_file_location(name, str(REF))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.CONFIG_PATH = Path('/nonexistent/.mtscomp')
    return m


@pytest.fixture(scope='module')
def plain():
    return _load_reference('mtscomp_reference_plain')


@pytest.fixture(scope='module')
def bound():
    from mtscomp_b200 import build
    sys.path.insert(0, str(ROOT / 'integration'))
    try:
        import reference_binding
    finally:
        sys.path.pop(0)
    m = _load_reference('mtscomp_reference_bound')
    m._b200 = reference_binding.install(m, build.build_emulation(), device=0)
    return m


def test_binding_declares_what_the_header_exports():
    """Every symbol the binding calls is declared in include/mtscomp_b200.h."""
    hdr = (ROOT / 'include' / 'mtscomp_b200.h').read_text()
    src = (ROOT / 'integration' / 'reference_binding.py').read_text()
    import re
    used = set(re.findall(r'lib\.(mtsb_\w+)', src))
    assert used >= {'mtsb_create', 'mtsb_compress_chunks', 'mtsb_decompress_chunks', 'mtsb_compress_bound'}
    for name in used:
        assert re.search(r'\b%s\s*\(' % name, hdr), name
    assert 'mtscomp_b200' not in re.sub(r'libmtscomp_b200|mtscomp_b200\.h', '', src.split('import ctypes')[1])


@pytest.mark.parametrize('kw', [dict(n_threads=1), dict(n_threads=3, do_spatial_diff=True), dict(chunk_order='C'),
                                dict(do_time_diff=False, do_spatial_diff=True), dict(chunk_duration=0.3)])
def test_reference_with_the_binding_round_trips_and_interoperates(plain, bound, tmp_path, kw):
    from mtscomp_b200 import synth
    arr = synth.ap_chunk(ns=2600, nc=50, sample_rate=30000., seed=11)
    raw = tmp_path / 'data.bin'
    arr.tofile(raw)
    common = dict(sample_rate=1000., n_channels=50, dtype='int16', quiet=True)
    # the reference's own compress() (its post-compression check included) with the codec behind the C ABI
    bound.compress(raw, tmp_path / 'b.cbin', tmp_path / 'b.ch', check_after_compress=True, **common, **kw)
    cd = bound._b200
    cd.lib.mtsb_last_launches.argtypes = [__import__('ctypes').c_void_p]
    assert cd.lib.mtsb_last_launches(cd.ctx) > 0          # (the kernels ran: the seam really goes through the library)
    # ... opens in the untouched reference
    r = plain.decompress(tmp_path / 'b.cbin', tmp_path / 'b.ch')
    assert np.array_equal(r[:], arr)
    r.close()
    # ... and in the bound one: whole array, slices within and across chunks (read_chunk + LRU, decompress_chunks), tofile
    g = bound.decompress(tmp_path / 'b.cbin', tmp_path / 'b.ch')
    assert np.array_equal(g[:], arr)
    assert np.array_equal(g[5:17], arr[5:17])
    assert np.array_equal(g[900:2100:7, 3:40], arr[900:2100:7, 3:40])
    assert np.array_equal(g[-1], arr[-1])
    g.close()
    bound.decompress(tmp_path / 'b.cbin', tmp_path / 'b.ch', tmp_path / 'b_back.bin', quiet=True,
                     check_after_decompress=True).close()
    assert (tmp_path / 'b_back.bin').read_bytes() == raw.read_bytes()
    # reference-written file through the binding
    plain.compress(raw, tmp_path / 'p.cbin', tmp_path / 'p.ch', **common, **kw)
    g = bound.decompress(tmp_path / 'p.cbin', tmp_path / 'p.ch')
    assert np.array_equal(g[:], arr)
    g.close()
    # same metadata apart from offsets and the digest of the compressed bytes; size within the north star's 3 %
    mb, mp = json.loads((tmp_path / 'b.ch').read_text()), json.loads((tmp_path / 'p.ch').read_text())
    assert set(mb) == set(mp)
    for k in mb:
        if k not in ('chunk_offsets', 'sha1_compressed'):
            assert mb[k] == mp[k], k
    assert (tmp_path / 'b.cbin').stat().st_size <= 1.031 * (tmp_path / 'p.cbin').stat().st_size + 64 * len(mb['chunk_bounds'])


def test_binding_reports_corruption_like_the_reference(plain, bound, tmp_path):
    from mtscomp_b200 import synth
    arr = synth.ap_chunk(ns=3000, nc=20, sample_rate=30000., seed=12)
    raw = tmp_path / 'data.bin'
    arr.tofile(raw)
    plain.compress(raw, tmp_path / 'p.cbin', tmp_path / 'p.ch', sample_rate=1000., n_channels=20, dtype='int16', quiet=True)
    ch = json.loads((tmp_path / 'p.ch').read_text())
    data = bytearray((tmp_path / 'p.cbin').read_bytes())
    mid = (ch['chunk_offsets'][1] + ch['chunk_offsets'][2]) // 2
    data[mid] ^= 0x5a
    (tmp_path / 'p.cbin').write_bytes(bytes(data))
    g = bound.Reader()
    g.open(tmp_path / 'p.cbin', tmp_path / 'p.ch')
    assert np.array_equal(g[:900], arr[:900])
    with pytest.raises(IOError, match=r'Compressed chunk #1 is corrupted\.'):
        g[1000:2000]
    g.close()
