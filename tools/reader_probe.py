#!/usr/bin/env python
"""Development probe for BASELINE configs[4]: random-access Reader slicing arr[t0:t1] through the drop-in Python API
(file on disk -> NumPy array), on a GPU-written and on a reference-format (.cbin of plain zlib streams) file.
usage: reader_probe.py [n_chunks]"""
import hashlib, json, sys, tempfile, time, zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import mtscomp_b200 as M
from mtscomp_b200 import synth
from oracle import codec as ora
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ns, nc = 30000, 385
tmp = Path(tempfile.mkdtemp())
M.CONFIG_PATH = tmp / '.mtscomp'
base = [np.ascontiguousarray(synth.ap_chunk(ns, nc, seed=300 + i)) for i in range(4)]
arr = np.concatenate([base[i % 4] for i in range(n_chunks)])
arr.tofile(tmp / 'data.bin')
t = time.perf_counter()
M.compress(tmp / 'data.bin', tmp / 'gpu.cbin', tmp / 'gpu.ch', sample_rate=30000., n_channels=nc, dtype=np.int16,
           check_after_compress=False, quiet=True)
print('Writer (file -> .cbin, %d chunks, %.2f GB): %.2f s' % (n_chunks, arr.nbytes / 1e9, time.perf_counter() - t))
# the same recording as the reference Writer would store it: plain zlib streams, same .ch with new offsets / checksum
ch = json.loads((tmp / 'gpu.ch').read_text())
with ThreadPoolExecutor(8) as ex:
    zs = list(ex.map(ora.encode_chunk, base))
streams = [zs[i % 4] for i in range(n_chunks)]
blob = b''.join(streams)
ch['chunk_offsets'] = [0] + [int(v) for v in np.cumsum([len(s) for s in streams])]
ch['sha1_compressed'] = hashlib.sha1(blob).hexdigest()
(tmp / 'ref.cbin').write_bytes(blob)
(tmp / 'ref.ch').write_text(json.dumps(ch, indent=2, sort_keys=True))
for name in ('gpu', 'ref'):
    r = M.decompress(tmp / (name + '.cbin'), tmp / (name + '.ch'))
    r[0:10]                                                         # warm-up (context, buffers)
    res = []
    for label, sl in (('100 samples inside one chunk', slice(45000, 45100)), ('one whole chunk', slice(60000, 90000)),
                      ('1.5 s across two chunks', slice(100000, 145000)), ('8 chunks', slice(90000, 330000)),
                      ('everything', slice(0, n_chunks * ns))):
        best = 1e9
        for _ in range(3):
            t = time.perf_counter(); x = r[sl]; best = min(best, time.perf_counter() - t)
        assert np.array_equal(x, arr[sl])
        res.append('%s: %.1f ms' % (label, best * 1e3))
    print('Reader on the %s-written file: ' % ('GPU' if name == 'gpu' else 'reference') + '; '.join(res))
    r.close()
