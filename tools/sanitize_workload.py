#!/usr/bin/env python
"""Small pass over every kernel of the codec, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
compress + decompress of AP chunks (default layout), LFP chunks (spatial diff), 'C' order, uint8 and int32 data, and the
decode of reference-written streams (block kernels, cells path, serial tails), each checked against the oracle."""
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, synth  # noqa: E402
from oracle import codec as ora  # noqa: E402

cd = _native.default_codec(0)


def roundtrip(x, rows, td=True, sd=False, order='F', label=''):
    fl = _native.flags_of(td, sd, order)
    comp, offs = cd.compress(x, rows, fl)
    for i in range(len(rows) - 1):
        assert zlib.decompress(bytes(comp[offs[i]:offs[i + 1]])) == ora.transform_chunk(x[rows[i]:rows[i + 1]], td, sd, order)
    out, st = cd.decompress(comp, offs, rows, x.shape[1], x.dtype, fl)
    assert not st.any() and np.array_equal(out, x), label
    parts = [ora.encode_chunk(x[rows[i]:rows[i + 1]], td, sd, order) for i in range(len(rows) - 1)]
    roffs = np.concatenate(([0], np.cumsum([len(p) for p in parts])))
    out, st = cd.decompress(b''.join(parts), roffs, rows, x.shape[1], x.dtype, fl)
    assert not st.any() and np.array_equal(out, x), label
    print('ok', label, flush=True)


roundtrip(synth.ap_chunk(6000, 385, seed=1), [0, 3000, 6000], label='ap 2 chunks')
roundtrip(synth.lfp_chunk(2500, 97, seed=2), [0, 1200, 2500], sd=True, label='lfp spatial')
roundtrip(synth.ap_chunk(1500, 33, seed=3), [0, 1500], order='C', label='order C')
roundtrip((synth.ap_chunk(20000, 8, seed=4) & 0xff).astype(np.uint8), [0, 20000], label='uint8')
roundtrip(synth.ap_chunk(4000, 16, seed=5).astype(np.int32), [0, 4000], label='int32')
roundtrip(synth.ap_chunk(30000, 24, seed=6), [0, 30000], label='ap long channels (multi-block streams)')
roundtrip(synth.ap_chunk(900, 7, seed=8).astype(np.int64), [0, 500, 900], label='int64 (look-back with value / tag arrays)')
roundtrip(synth.ap_chunk(400, 900, seed=9), [0, 150, 400], label='900 channels (two per thread)')
roundtrip(synth.ap_chunk(300, 1800, seed=10), [0, 300], sd=True, label='1800 channels (four per thread), spatial')
roundtrip(synth.ap_chunk(200, 2100, seed=11), [0, 200], label='2100 channels (generic transform kernels)')
for cells in (1, 0):
    cd.set_param('par_cells', cells)
    roundtrip(synth.ap_chunk(30000, 16, seed=7), [0, 30000], label='reference streams, par_cells=%d' % cells)
cd.set_param('par_cells', -1)
print('sanitize workload done')
