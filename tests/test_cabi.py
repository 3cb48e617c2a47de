"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/mtscomp_b200.h declares."""
import ctypes
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def lib_path():
    from mtscomp_b200 import build
    return build.build_native()


def declared_symbols():
    text = (ROOT / 'include' / 'mtscomp_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mtsb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_binding_covers_header():
    from mtscomp_b200 import _native
    assert sorted(_native.SYMBOLS) == declared_symbols()


def test_library_is_sm100a(lib_path):
    import subprocess
    out = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-lelf', str(lib_path)], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_no_cpu_fallback_without_device(lib_path):
    """On a box without a GPU the product path must fail loudly, not fall back."""
    from mtscomp_b200 import _native
    lib = _native.load_library()
    if lib.mtsb_device_count() > 0:
        pytest.skip('a CUDA device is present')
    with pytest.raises(_native.NativeUnavailable):
        _native.Codec(0)
    import numpy as np
    import mtscomp_b200
    with pytest.raises(_native.NativeUnavailable):
        mtscomp_b200.diff_along_axis(np.zeros((4, 4), np.int16), 0)


def test_product_never_imports_oracle():
    import pathlib
    for p in (ROOT / 'mtscomp_b200').rglob('*.py'):
        src = p.read_text()
        assert 'import oracle' not in src and 'from oracle' not in src, p
        assert 'import zlib' not in src, p
