"""CPU: the same drop-in API cases with the kernels' logic run through the host emulation (development aid; the
package itself never selects it — the fixture injects the emulated context where the CUDA one would be)."""
import pytest

import api_cases


@pytest.fixture(scope='module')
def emulated_default_codec():
    from mtscomp_b200 import _native, build
    emu = _native.Codec(0, lib=_native.load_library(build.build_emulation()))
    saved = dict(_native._default)
    _native._default.clear()
    _native._default[0] = emu
    yield emu
    _native._default.clear()
    _native._default.update(saved)


@pytest.mark.parametrize('case', api_cases.ALL_CASES, ids=lambda f: f.__name__)
def test_api_emulated(case, tmp_path, emulated_default_codec):
    case(tmp_path)
