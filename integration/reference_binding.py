# -*- coding: utf-8 -*-
"""The reference-side binding of the C ABI, as a file a maintainer of int-brain-lab/mtscomp could
drop next to `mtscomp.py`: ctypes only, no import of this package.

    import mtscomp, reference_binding
    reference_binding.install(mtscomp, '/path/to/libmtscomp_b200.so', device=0)
    mtscomp.compress(...); mtscomp.decompress(...)        # the reference's own code, codec on the B200

`install` replaces the three methods that make up the reference's codec seam and nothing else:

  Writer.compress_batch    (mtscomp.py:399-423; fans out _compress_chunk, :375-397)   -> mtsb_compress_chunks
  Reader.decompress_chunks (mtscomp.py:645-650; fans out read_chunk, :602-635)         -> mtsb_decompress_chunks
  Reader.read_chunk        (mtscomp.py:602-635; one chunk, wrapped in the LRU at :582-588) -> mtsb_decompress_chunks

File handling, offsets, SHA-1 digests, the `.ch` metadata, slicing, the cache and the CLI stay the reference's.
There is no CPU fallback: if the library cannot be loaded or a call fails, the error is raised.
tests/test_reference_binding.py runs the unmodified reference with this binding installed."""
import ctypes as C
import os
import threading

import numpy as np

_PLL = C.POINTER(C.c_longlong)
MTSB_E_CORRUPT = -5
TIME_DIFF, SPATIAL_DIFF, ORDER_C, FLOAT = 1, 2, 4, 8


def _declare(lib):
    vp, ci, ll = C.c_void_p, C.c_int, C.c_longlong
    lib.mtsb_create.restype = vp
    lib.mtsb_create.argtypes = [ci, vp]
    lib.mtsb_destroy.argtypes = [vp]
    lib.mtsb_last_error.restype = C.c_char_p
    lib.mtsb_last_error.argtypes = [vp]
    lib.mtsb_compress_bound.restype = ll
    lib.mtsb_compress_bound.argtypes = [vp, ll, ll, ci, ci, ci]
    lib.mtsb_compress_chunks.argtypes = [vp, vp, ci, ci, _PLL, ci, ci, ci, vp, ci, ll, _PLL]
    lib.mtsb_decompress_chunks.argtypes = [vp, vp, ci, _PLL, ci, _PLL, ci, ci, ci, vp, ci, C.POINTER(ci)]
    return lib


class _Codec:
    """One context (one CUDA stream and its scratch); calls are serialised, as include/mtscomp_b200.h requires."""

    def __init__(self, lib_path, device):
        self.lib = _declare(C.CDLL(str(lib_path)))
        self.ctx = self.lib.mtsb_create(int(device), None)
        if not self.ctx:
            raise RuntimeError('mtsb_create failed: %s' % self.lib.mtsb_last_error(None).decode())
        self.lock = threading.Lock()

    def error(self):
        return self.lib.mtsb_last_error(self.ctx).decode()


def _flags(do_time_diff, do_spatial_diff, chunk_order, dtype):
    return ((TIME_DIFF if do_time_diff else 0) | (SPATIAL_DIFF if do_spatial_diff else 0) |
            (ORDER_C if chunk_order == 'C' else 0) | (FLOAT if np.dtype(dtype).kind == 'f' else 0))


def install(mtscomp, lib_path, device=0):
    """Route the codec seam of the reference module `mtscomp` through the shared library at `lib_path`."""
    cd = _Codec(lib_path, device)

    def compress_batch(self, first_chunk, last_chunk):
        assert 0 <= first_chunk < last_chunk <= self.n_chunks
        b = self.chunk_bounds
        n = last_chunk - first_chunk
        # the batch's rows, contiguous in host memory (chunks of a file are consecutive rows of the memory map)
        block = np.ascontiguousarray(self.data[b[first_chunk]:b[last_chunk]])
        assert block.ndim == 2 and block.shape[1] == self.n_channels
        rows = np.asarray(b[first_chunk:last_chunk + 1], dtype=np.int64) - b[first_chunk]
        nc, isz = block.shape[1], block.itemsize
        fl = _flags(self.do_time_diff, self.do_spatial_diff, self.chunk_order, block.dtype)
        with cd.lock:
            cap = sum(cd.lib.mtsb_compress_bound(cd.ctx, int(rows[i + 1] - rows[i]) * nc * isz, int(rows[i + 1] - rows[i]),
                                                 nc, isz, fl) for i in range(n))
            dst = np.empty(cap, dtype=np.uint8)
            offs = np.zeros(n + 1, dtype=np.int64)
            rc = cd.lib.mtsb_compress_chunks(cd.ctx, block.ctypes.data, 0, n, rows.ctypes.data_as(_PLL), nc, isz, fl,
                                             dst.ctypes.data, 0, cap, offs.ctypes.data_as(_PLL))
            if rc != 0:
                raise RuntimeError('mtsb_compress_chunks: %s' % cd.error())
        return {first_chunk + i: (block[rows[i]:rows[i + 1]], dst[offs[i]:offs[i + 1]].tobytes()) for i in range(n)}

    def _decode(self, ids):
        if not ids:
            return {}
        fd = self.cdata.fileno()
        bufs = [os.pread(fd, self.chunk_offsets[i + 1] - self.chunk_offsets[i], self.chunk_offsets[i]) for i in ids]
        comp = np.frombuffer(b''.join(bufs), dtype=np.uint8)
        offs = np.concatenate(([0], np.cumsum([len(x) for x in bufs]))).astype(np.int64)
        rows = np.concatenate(([0], np.cumsum([self.chunk_bounds[i + 1] - self.chunk_bounds[i] for i in ids]))).astype(np.int64)
        out = np.empty((int(rows[-1]), self.n_channels), dtype=self.dtype)
        st = np.zeros(len(ids), dtype=np.int32)
        fl = _flags(self.cmeta.do_time_diff, self.cmeta.do_spatial_diff, self.chunk_order, self.dtype)
        with cd.lock:
            rc = cd.lib.mtsb_decompress_chunks(cd.ctx, comp.ctypes.data, 0, offs.ctypes.data_as(_PLL), len(ids),
                                               rows.ctypes.data_as(_PLL), self.n_channels, self.dtype.itemsize, fl,
                                               out.ctypes.data, 0, st.ctypes.data_as(C.POINTER(C.c_int)))
            if rc == MTSB_E_CORRUPT:          # the reference's own error (mtscomp.py:621)
                raise IOError("Compressed chunk #%d is corrupted." % ids[int(np.flatnonzero(st)[0])])
            if rc != 0:
                raise RuntimeError('mtsb_decompress_chunks: %s' % cd.error())
        return {i: out[rows[k]:rows[k + 1]] for k, i in enumerate(ids)}

    def decompress_chunks(self, chunk_ids, pool=None):
        ids = list(chunk_ids)
        out = _decode(self, ids)
        assert set(out.keys()) == set(ids)
        return out

    def read_chunk(self, chunk_idx, chunk_start, chunk_length):
        assert chunk_start == self.chunk_offsets[chunk_idx]
        assert chunk_length == self.chunk_offsets[chunk_idx + 1] - chunk_start
        return np.ascontiguousarray(_decode(self, [chunk_idx])[chunk_idx])

    mtscomp.Writer.compress_batch = compress_batch
    mtscomp.Reader.decompress_chunks = decompress_chunks
    mtscomp.Reader.read_chunk = read_chunk
    return cd
