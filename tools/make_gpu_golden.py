#!/usr/bin/env python
"""Run ON THE B200: write small recordings with this package's Writer (CUDA library, no emulation) into
gpurun_out/gpu_golden/; the files are then committed under tests/golden/gpu_written/ and opened by the UNMODIFIED
reference Reader in the CPU suite (tests/test_reference_interop.py)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import mtscomp_b200 as M  # noqa: E402
from mtscomp_b200 import _native, synth  # noqa: E402

out = ROOT / 'gpurun_out' / 'gpu_golden'
out.mkdir(parents=True, exist_ok=True)
M.CONFIG_PATH = out / '.mtscomp'
cd = _native.default_codec(0)
assert str(_native.LIB_PATH).endswith('libmtscomp_b200.so') and cd.get_param('sm_count') > 100
cases = {
    'ap_int16': (synth.ap_chunk(ns=2600, nc=50, sample_rate=30000., seed=3), dict()),
    'lfp_spatial': (synth.lfp_chunk(ns=2500, nc=40, seed=8), dict(do_spatial_diff=True)),
    'ap_order_c_short_chunks': (synth.ap_chunk(ns=1500, nc=20, sample_rate=30000., seed=4), dict(chunk_order='C', chunk_duration=0.4)),
}
manifest = {}
for name, (arr, kw) in cases.items():
    arr.tofile(out / (name + '.bin'))
    M.compress(out / (name + '.bin'), out / (name + '.cbin'), out / (name + '.ch'), sample_rate=1000., n_channels=arr.shape[1],
               dtype='int16', quiet=True, **kw)
    r = M.decompress(out / (name + '.cbin'), out / (name + '.ch'))
    assert np.array_equal(r[:], arr)
    r.close()
    manifest[name] = {'shape': list(arr.shape), 'dtype': 'int16', 'kwargs': kw, 'device': 'B200 (sm_count %d)' % cd.get_param('sm_count')}
(out / 'manifest.json').write_text(json.dumps(manifest, indent=1, sort_keys=True))
(out / '.mtscomp').unlink(missing_ok=True)
print('wrote', sorted(p.name for p in out.iterdir()))
