#!/usr/bin/env python
"""Host-only probe: ways to fill a fresh tmpfs file from one large buffer (what Reader.tofile does per batch)."""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

n = 1 << 30
src = np.random.default_rng(0).integers(0, 255, n, dtype=np.uint8)
p = '/dev/shm/_mtsb_wtest.bin'


def t_write():
    with open(p, 'wb') as f:
        f.write(src)


def t_pwrite(k):
    with open(p, 'wb') as f, ThreadPoolExecutor(k) as ex:
        fd, mv, st = f.fileno(), memoryview(src), n // k
        list(ex.map(lambda a: os.pwrite(fd, mv[a:a + st], a), range(0, n, st)))


def t_mmap(k):
    with open(p, 'w+b') as f, ThreadPoolExecutor(k) as ex:
        f.truncate(n)
        mm = np.memmap(f, dtype=np.uint8, mode='r+', shape=(n,))
        st = n // k
        list(ex.map(lambda a: np.copyto(mm[a:a + st], src[a:a + st]), range(0, n, st)))
        del mm


for name, fn in (('write', t_write), ('pwrite x4', lambda: t_pwrite(4)), ('mmap x4', lambda: t_mmap(4)),
                 ('mmap x8', lambda: t_mmap(8)), ('write again', t_write)):
    if os.path.exists(p):
        os.unlink(p)
    t = time.perf_counter()
    fn()
    print('%-12s %.2f GB/s' % (name, n / (time.perf_counter() - t) / 1e9), flush=True)
os.unlink(p)
