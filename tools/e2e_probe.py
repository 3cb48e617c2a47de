"""Development probe: does the host-buffer path overlap copies with kernels?"""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mtscomp_b200 import _native, synth
n_chunks = 128; ns, nc = 30000, 385
cd = _native.default_codec(0)
base = [synth.ap_chunk(ns, nc, seed=100 + i) for i in range(4)]
raw_bytes = n_chunks * ns * nc * 2
h_raw = torch.empty(raw_bytes, dtype=torch.uint8, pin_memory=True)
hv = h_raw.numpy()
for i in range(n_chunks): hv[i * ns * nc * 2:(i + 1) * ns * nc * 2] = base[i % 4].reshape(-1).view(np.uint8)
rows = np.arange(n_chunks + 1, dtype=np.int64) * ns
cap = n_chunks * cd.compress_bound(ns, nc, 2, 1)
h_comp = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
d_raw = h_raw.cuda(); d_comp = torch.empty(cap, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize()
for hb in (256 << 20, 8 << 30):
    cd.set_param('host_batch_bytes', hb); cd.set_param('batch_bytes', max(hb, 2 << 30))
    for rep in range(2):
        t = time.perf_counter(); offs = cd.compress_ptr(h_raw.data_ptr(), 0, rows, nc, 2, 1, h_comp.data_ptr(), 0, cap); dt = time.perf_counter() - t
    tm = cd.timings()
    t = time.perf_counter(); cd.compress_ptr(d_raw.data_ptr(), 1, rows, nc, 2, 1, d_comp.data_ptr(), 1, cap); dd = time.perf_counter() - t
    print('host_batch %5d MB: e2e %.1f ms (%.1f GB/s)   device-resident %.1f ms   pure copies would take %.1f ms' % (
        hb >> 20, dt * 1e3, raw_bytes / dt / 1e9, dd * 1e3, (raw_bytes + int(offs[-1])) / 55e9 * 1e3), ' stages', ['%.1f' % v for v in tm], flush=True)

# ---- does a plain H2D copy make progress while the persistent lz77 kernel owns every SM?
import threading
cd.set_param('batch_bytes', 8 << 30)
n = 2 << 30
hh = torch.empty(n, dtype=torch.uint8, pin_memory=True); dd_ = torch.empty(n, dtype=torch.uint8, device='cuda')
s2 = torch.cuda.Stream()
def copy_only():
    torch.cuda.synchronize(); t = time.perf_counter()
    with torch.cuda.stream(s2): dd_.copy_(hh, non_blocking=True)
    s2.synchronize(); return (time.perf_counter() - t) * 1e3
print('H2D 2 GiB alone: %.1f ms' % copy_only(), flush=True)
res = {}
def comp():
    t = time.perf_counter(); cd.compress_ptr(d_raw.data_ptr(), 1, rows, nc, 2, 1, d_comp.data_ptr(), 1, cap); res['c'] = (time.perf_counter() - t) * 1e3
th = threading.Thread(target=comp); th.start(); time.sleep(0.005)
t = time.perf_counter()
with torch.cuda.stream(s2): dd_.copy_(hh, non_blocking=True)
s2.synchronize(); res['h'] = (time.perf_counter() - t) * 1e3
th.join()
print('concurrent: compress %.1f ms, H2D 2 GiB %.1f ms' % (res['c'], res['h']), flush=True)
