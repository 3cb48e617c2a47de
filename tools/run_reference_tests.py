#!/usr/bin/env python
"""Run the REFERENCE's own test-suite (/root/reference/tests.py, unmodified, 192 tests) against mtscomp_b200.

Build container only (the reference tree does not travel).  `import mtscomp` inside the suite is answered with this
package through an import alias; on a box without a GPU the kernels' logic runs through the host emulation
(csrc/emu, development aid), exactly as tests/test_emu_api.py does — pass --gpu to use the CUDA library instead.
With --binding the suite's `import mtscomp` gets the UNMODIFIED reference module with integration/reference_binding.py
installed: the reference's own host code, only its codec seam routed through the C ABI.
The summary goes to stdout (committed under profiles/ per round)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF_TESTS = Path('/root/reference/tests.py')


def main():
    import pytest  # noqa: F401
    import mtscomp_b200
    from mtscomp_b200 import _native, build
    if '--binding' in sys.argv:
        import importlib.util
        sys.path.insert(0, str(ROOT / 'integration'))
        import reference_binding
        spec = importlib.util.spec_from_file_location('mtscomp', '/root/reference/mtscomp.py')
        ref = importlib.util.module_from_spec(spec)
        sys.modules['mtscomp'] = ref
        spec.loader.exec_module(ref)
        reference_binding.install(ref, build.build_native() if '--gpu' in sys.argv else build.build_emulation(), 0)
        return pytest.main(_args())
    if '--gpu' not in sys.argv:
        emu = _native.Codec(0, lib=_native.load_library(build.build_emulation()))
        _native._default.clear()
        _native._default[0] = emu
    sys.modules['mtscomp'] = mtscomp_b200          # the alias: the suite's `import mtscomp` gets this package
    return pytest.main(_args())


def _args():
    args = [str(REF_TESTS), '-q', '-x' if '-x' in sys.argv else '--maxfail=1000', '-p', 'no:cacheprovider',
            '--rootdir', '/tmp', '-o', 'python_files=tests.py']
    if '-k' in sys.argv:
        args += ['-k', sys.argv[sys.argv.index('-k') + 1]]
    return args


if __name__ == '__main__':
    sys.exit(main())
