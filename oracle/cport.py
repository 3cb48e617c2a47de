# -*- coding: utf-8 -*-
"""ORACLE (test infrastructure): ctypes access to oracle/c/mtsoracle.c, the dependency-free C restatement."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / '_build' / 'libmtsoracle.so'


def build():
    src = _DIR / 'c' / 'mtsoracle.c'
    if not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(['make', '-C', str(_DIR)], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.ora_inflate.restype = C.c_long
        _lib.ora_inflate.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        _lib.ora_adler32.restype = C.c_uint32
        _lib.ora_adler32.argtypes = [C.c_void_p, C.c_long]
        for f in (_lib.ora_transform, _lib.ora_untransform):
            f.restype = None
            f.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int, C.c_void_p]
    return _lib


def _flags(td, sd, order):
    return (1 if td else 0) | (2 if sd else 0) | (4 if order == 'C' else 0)


def transform(chunk, td=True, sd=False, order='F'):
    chunk = np.ascontiguousarray(chunk)
    out = np.empty(chunk.nbytes, np.uint8)
    lib().ora_transform(chunk.ctypes.data, chunk.shape[0], chunk.shape[1], chunk.dtype.itemsize, _flags(td, sd, order),
                        out.ctypes.data)
    return out.tobytes()


def untransform(buf, ns, nc, dtype, td=True, sd=False, order='F'):
    src = np.frombuffer(buf, np.uint8)
    out = np.empty((ns, nc), dtype)
    lib().ora_untransform(src.ctypes.data, ns, nc, out.dtype.itemsize, _flags(td, sd, order), out.ctypes.data)
    return out


def inflate(cbuf, out_len):
    src = np.frombuffer(bytes(cbuf), np.uint8)
    out = np.empty(max(out_len, 1), np.uint8)
    n = lib().ora_inflate(src.ctypes.data, len(src), out.ctypes.data, out_len)
    if n < 0:
        raise ValueError('ora_inflate error %d' % n)
    return out[:n].tobytes()


def adler32(buf):
    src = np.frombuffer(bytes(buf), np.uint8)
    return int(lib().ora_adler32(src.ctypes.data, len(src)))
