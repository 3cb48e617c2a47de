#!/usr/bin/env python
"""DEVELOPMENT TOOL: compression ratio of the encoder (kernel logic via host emulation) vs zlib on synthetic data.
usage: emu_ratio.py ap|lfp [param=value ...]"""
import sys, zlib, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tools'))
import emu_check
from mtscomp_b200 import synth, _native
from oracle import codec as ora

kind = sys.argv[1]
cd = emu_check.get_codec()
for kv in sys.argv[2:]:
    k, v = kv.split('=')
    cd.set_param(k, int(v))
if kind == 'ap':
    x = np.ascontiguousarray(synth.ap_chunk(30000, 385, seed=1234)[:, 100:132]); sd = False   # 8 segments of 4 channels
else:
    x = synth.lfp_chunk(2500, 385, seed=50); sd = True
fl = _native.flags_of(True, sd, 'F')
t = time.time()
comp, offs = cd.compress(x, [0, x.shape[0]], fl)
want = ora.transform_chunk(x, True, sd)
assert zlib.decompress(bytes(comp)) == want
ref = len(zlib.compress(want))
print('%s %s size/zlib %.4f  (%d vs %d) %.0fs' % (kind, ' '.join(sys.argv[2:]), len(comp) / ref, len(comp), ref, time.time() - t))
