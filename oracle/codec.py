# -*- coding: utf-8 -*-
"""ORACLE (test infrastructure, not product code): NumPy + zlib restatement of mtscomp's per-chunk codec.

Parity status: PINNED.  tests/test_oracle.py checks every function here against golden vectors produced by the
unmodified reference (`/root/reference/mtscomp.py`, imported by tools/make_golden.py in the build container):
byte-identical `.cbin` payloads when the runtime zlib equals the one that wrote the fixtures (zlib 1.3), and
identical decoded arrays always.

Third-party arithmetic on this path that is NOT in the reference tree: DEFLATE + adler32 from system zlib, reached
through CPython's `zlib` module (reference call sites mtscomp.py:394 `zlib.compress(bytes)` -> level 6 / wbits 15 /
memLevel 8, and mtscomp.py:619 `zlib.decompress`).  The reference pins no version (requirements.txt:1); this image
has zlib 1.3.  The oracle calls the same module, so the DEFLATE bit-stream semantics are RFC 1950/1951 exactly as the
reference sees them; oracle/c/mtsoracle.c restates inflate/adler32 independently for cross-checking.
"""

from concurrent.futures import ThreadPoolExecutor
import zlib

import numpy as np


def diff_along_axis(chunk, axis=None):
    """mtscomp.py:143-159 — np.diff along `axis`, first row (axis 0) / first column (axis 1) kept.  Integer
    arithmetic wraps modulo 2**(8*itemsize)."""
    if axis is None:
        return chunk
    d = np.empty_like(chunk)
    if axis == 0:
        d[:1, :] = chunk[:1, :]
        np.subtract(chunk[1:, :], chunk[:-1, :], out=d[1:, :])
    else:
        d[:, :1] = chunk[:, :1]
        np.subtract(chunk[:, 1:], chunk[:, :-1], out=d[:, 1:])
    return d


def cumsum_along_axis(chunk, axis=None):
    """mtscomp.py:162-169 — np.cumsum(..., out=empty_like(chunk)): result dtype == input dtype, wraps."""
    if axis is None:
        return chunk
    out = np.empty_like(chunk)
    np.cumsum(chunk, axis=axis, out=out)
    return out


def transform_chunk(chunk, do_time_diff=True, do_spatial_diff=False, chunk_order='F'):
    """mtscomp.py:381-394 up to (not including) zlib.compress: the exact bytes handed to deflate."""
    d = diff_along_axis(chunk, 0 if do_time_diff else None)
    d = diff_along_axis(d, 1 if do_spatial_diff else None)
    return d.tobytes(order=chunk_order)


def encode_chunk(chunk, do_time_diff=True, do_spatial_diff=False, chunk_order='F'):
    """mtscomp.py:375-397 `Writer._compress_chunk`: transform + zlib.compress at the default level
    (comp_level is never forwarded)."""
    return zlib.compress(transform_chunk(chunk, do_time_diff, do_spatial_diff, chunk_order))


def untransform_bytes(buf, n_samples, n_channels, dtype, do_time_diff=True, do_spatial_diff=False, chunk_order='F'):
    """mtscomp.py:622-635: frombuffer -> reshape(order) -> cumsum(space) -> cumsum(time) -> C-contiguous."""
    a = np.frombuffer(buf, dtype=dtype)
    assert a.size == n_samples * n_channels
    a = a.reshape((n_samples, n_channels), order=chunk_order)
    a = cumsum_along_axis(a, 1 if do_spatial_diff else None)
    a = cumsum_along_axis(a, 0 if do_time_diff else None)
    return np.ascontiguousarray(a)


def decode_chunk(cbuf, n_samples, n_channels, dtype, do_time_diff=True, do_spatial_diff=False, chunk_order='F'):
    """mtscomp.py:602-635 `Reader.read_chunk` minus the pread: zlib.decompress + inverse transform.
    Any zlib failure (bad header, bad adler32, truncated) propagates, as the reference turns it into IOError."""
    return untransform_bytes(zlib.decompress(cbuf), n_samples, n_channels, dtype,
                             do_time_diff, do_spatial_diff, chunk_order)


def chunk_bounds(n_samples, sample_rate, chunk_duration):
    """mtscomp.py:324-335."""
    cs = int(np.round(chunk_duration * sample_rate))
    b = list(range(0, n_samples, cs))
    if b[-1] < n_samples:
        b.append(n_samples)
    return b


def encode_array(data, sample_rate, chunk_duration=1.0, n_threads=1, **flags):
    """mtscomp.py:425-489 (`Writer.write` loop) without files: returns (cbin_bytes, chunk_bounds, chunk_offsets).
    Threads mirror the reference's ThreadPool batches (`mtscomp.py:456-469`): zlib releases the GIL."""
    bounds = chunk_bounds(data.shape[0], sample_rate, chunk_duration)
    ids = range(len(bounds) - 1)

    def one(i):
        return encode_chunk(data[bounds[i]:bounds[i + 1]], **flags)
    if n_threads <= 1:
        parts = [one(i) for i in ids]
    else:
        with ThreadPoolExecutor(n_threads) as ex:
            parts = list(ex.map(one, ids))
    offsets = [0]
    for p in parts:
        offsets.append(offsets[-1] + len(p))
    return b''.join(parts), bounds, offsets


def decode_array(cbin, bounds, offsets, n_channels, dtype, n_threads=1, **flags):
    """mtscomp.py:701-735 (`Reader.tofile` loop) without files: returns the (n_samples, n_channels) array."""
    ids = range(len(bounds) - 1)
    mv = memoryview(cbin)

    def one(i):
        return decode_chunk(mv[offsets[i]:offsets[i + 1]], bounds[i + 1] - bounds[i], n_channels, dtype, **flags)
    if n_threads <= 1:
        parts = [one(i) for i in ids]
    else:
        with ThreadPoolExecutor(n_threads) as ex:
            parts = list(ex.map(one, ids))
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, n_channels), dtype)


def adler32_combine(a1, a2, len2):
    """zlib's adler32_combine, used to check the GPU's per-sub-block adler stitching."""
    BASE = 65521
    s1a, s2a = a1 & 0xffff, (a1 >> 16) & 0xffff
    s1b, s2b = a2 & 0xffff, (a2 >> 16) & 0xffff
    s1 = (s1a + s1b - 1) % BASE
    s2 = (s2a + s2b + (len2 % BASE) * (s1a - 1)) % BASE
    return (s2 << 16) | s1
