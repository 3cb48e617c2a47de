"""GPU: the drop-in Writer / Reader / compress / decompress / CLI behaviours (tests/api_cases.py) on the B200."""
import pytest

import api_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', api_cases.ALL_CASES, ids=lambda f: f.__name__)
def test_api(case, tmp_path, codec):
    case(tmp_path)
