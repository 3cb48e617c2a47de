import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def codec():
    """The product's native codec on cuda:0.  No fallback: GPU tests fail loudly if it cannot be created."""
    from mtscomp_b200 import _native
    return _native.default_codec(0)


GOLDEN = ROOT / 'tests' / 'golden'
