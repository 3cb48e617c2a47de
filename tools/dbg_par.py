"""Development probe: decode n reference-written chunks (device path through host arrays), print what the
block-parallel decoder did and its stage times.  usage: dbg_par.py [n_chunks] [param=value ...]"""
import sys, time, zlib
sys.path.insert(0, '/root/repo')
import numpy as np
from mtscomp_b200 import _native, synth
from oracle import codec as ora
cd = _native.default_codec(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for kv in sys.argv[2:]:
    k, v = kv.split('=')
    cd.set_param(k, int(v))
base = [np.ascontiguousarray(synth.ap_chunk(30000, 385, seed=100 + i)) for i in range(min(n, 8))]
zs = [ora.encode_chunk(b) for b in base]
comp = b''.join(zs[i % len(zs)] for i in range(n))
offs = np.zeros(n + 1, dtype=np.int64); offs[1:] = np.cumsum([len(zs[i % len(zs)]) for i in range(n)])
rows = np.arange(n + 1, dtype=np.int64) * 30000
for rep in range(3):
    t = time.perf_counter()
    out, st = cd.decompress(comp, offs, rows, 385, np.int16, _native.TIME_DIFF)
    dt = time.perf_counter() - t
    print('n', n, 'ms %.1f' % (dt * 1e3), 'status', int(st.any()), 'cands', cd.get_param('par_candidates'), 'chained', cd.get_param('par_chained'), 'resumed', cd.get_param('par_resumed'), ['%.1f' % v for v in cd.timings()])
print('exact', all(np.array_equal(out[i * 30000:(i + 1) * 30000], base[i % len(base)]) for i in range(n)))
