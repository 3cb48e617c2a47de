#!/usr/bin/env python
"""Development: the K4 stage (inv_tile_kernel) of a device-resident decode under its tile-shape settings.

    python tools/inv_variants.py [n_chunks] [n_distinct]

Compresses n_chunks AP chunks once on the GPU, then decodes them with inv_persistent = 0 / 1 and a few
inv_order_block values and prints the `inverse` stage time of each (CUDA events inside the library) next to the HBM
fraction it amounts to (2 bytes moved per raw byte)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, synth  # noqa: E402


def main():
    n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    n_distinct = min(n_chunks, int(sys.argv[2]) if len(sys.argv) > 2 else 4)
    ns, nc = 30000, 385
    peak = 6540.5
    try:
        peak = float(json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()).get('hbm_gbs', peak))
    except Exception:
        pass
    cd = _native.default_codec(0)
    base = [synth.ap_chunk(ns, nc, seed=100 + i) for i in range(n_distinct)]
    x = np.ascontiguousarray(np.concatenate([base[i % n_distinct] for i in range(n_chunks)], axis=0))
    rows = np.arange(n_chunks + 1, dtype=np.int64) * ns
    fl = _native.TIME_DIFF
    lib = cd.lib
    raw_bytes = x.nbytes
    d_raw = lib.mtsb_device_alloc(cd.ctx, raw_bytes)
    cap = sum(cd.compress_bound(ns, nc, 2, fl) for _ in range(n_chunks))
    d_comp = lib.mtsb_device_alloc(cd.ctx, cap)
    d_out = lib.mtsb_device_alloc(cd.ctx, raw_bytes)
    lib.mtsb_memcpy(cd.ctx, d_raw, x.ctypes.data, raw_bytes, 1)
    offs = cd.compress_ptr(d_raw, 1, rows, nc, 2, fl, d_comp, 1, cap)
    back = np.empty_like(x)
    for half, ob in ((0, 2), (1, 2), (1, 1), (1, 4), (0, 1), (1, 2), (0, 2)):
        cd.set_param('inv_persistent', half)
        cd.set_param('inv_order_block', ob)
        best = None
        for it in range(3):
            t = time.time()
            st = cd.decompress_ptr(d_comp, 1, offs, rows, nc, 2, fl, d_out, 1)
            dt = time.time() - t
            tm = cd.timings()
            if best is None or tm[4] < best[0]:
                best = (tm[4], tm[7], dt)
        assert not st.any()
        print('inv_persistent=%d inv_order_block=%d  inverse %.3f ms = %.3f of HBM peak  (decode total %.2f ms, %.1f GB/s)' % (
            half, ob, best[0], 2 * raw_bytes / (best[0] * 1e-3) / 1e9 / peak, best[1], raw_bytes / (best[1] * 1e-3) / 1e9),
            flush=True)
        if half:
            lib.mtsb_memcpy(cd.ctx, back.ctypes.data, d_out, min(raw_bytes, 40 * ns * nc * 2), 2)
            assert np.array_equal(back[:40 * ns], x[:40 * ns])
    print('exact', flush=True)


if __name__ == '__main__':
    main()
