# -*- coding: utf-8 -*-
"""ORACLE (test infrastructure, not product code): CPU restatement of the reference Reader's random-access path.

Follows mtscomp.py:582-588 (LRU of decoded chunks, `cache_size` entries), :602-635 (`read_chunk`: pread + zlib +
inverse transform), :661-684 (`_chunks_for_interval`) and :798-856 (`__getitem__` for slices: whole chunks decoded one
by one, concatenated, then cut).  Used as the CPU baseline of the latency measurement (bench.py, BASELINE configs[4])
and as a checker in tests; pinned against the unmodified reference by tests/test_reference_interop.py.
"""
import bisect
import json
import os
from collections import OrderedDict

import numpy as np

from . import codec as ora


class PortReader:
    def __init__(self, cbin_path, ch_path, cache_size=10):
        with open(ch_path) as f:
            self.meta = json.load(f)
        self.fd = os.open(cbin_path, os.O_RDONLY)
        m = self.meta
        self.bounds, self.offsets = m['chunk_bounds'], m['chunk_offsets']
        self.n_channels, self.dtype = m['n_channels'], np.dtype(m['dtype'])
        self.n_samples = self.bounds[-1]
        self.n_chunks = len(self.bounds) - 1
        self.flags = dict(do_time_diff=m['do_time_diff'], do_spatial_diff=m['do_spatial_diff'], chunk_order=m['chunk_order'])
        self.cache_size = cache_size
        self.cache = OrderedDict()

    def close(self):
        os.close(self.fd)

    def read_chunk(self, idx):                                   # mtscomp.py:602-635 under the LRU of :582-588
        if idx in self.cache:
            self.cache.move_to_end(idx)
            return self.cache[idx]
        buf = os.pread(self.fd, self.offsets[idx + 1] - self.offsets[idx], self.offsets[idx])
        a = ora.decode_chunk(buf, self.bounds[idx + 1] - self.bounds[idx], self.n_channels, self.dtype, **self.flags)
        self.cache[idx] = a
        while len(self.cache) > self.cache_size:
            self.cache.popitem(last=False)
        return a

    def chunks_for_interval(self, i0, i1):                       # mtscomp.py:661-684
        clip = lambda x, lo, hi: max(lo, min(hi, x))
        i0 = clip(i0, 0, self.n_samples - 1)
        i1 = clip(i1, i0, self.n_samples - 1)
        first = clip(bisect.bisect_right(self.bounds, i0) - 1, 0, self.n_chunks - 1)
        last = clip(bisect.bisect_right(self.bounds, i1, lo=first) - 1, 0, self.n_chunks - 1)
        return first, last

    def __getitem__(self, item):                                 # mtscomp.py:798-856: slices, (slice, columns)
        if isinstance(item, tuple):
            return self[item[0]] if len(item) == 1 else self[item[0]][:, item[1]]
        assert isinstance(item, slice)
        i0 = 0 if item.start is None else int(item.start)
        i1 = self.n_samples if item.stop is None else int(item.stop)
        i0 = max(0, min(self.n_samples, i0 + self.n_samples if i0 < 0 else i0))
        i1 = max(0, min(self.n_samples, i1 + self.n_samples if i1 < 0 else i1))
        if i1 <= i0:
            return np.zeros((0, self.n_channels), dtype=self.dtype)
        first, last = self.chunks_for_interval(i0, i1)
        chunks = [self.read_chunk(i) for i in range(first, last + 1)]
        arr = chunks[0] if len(chunks) == 1 else np.concatenate(chunks)
        a, b = i0 - self.bounds[first], i1 - self.bounds[first]
        return arr[a:b:item.step, :]
