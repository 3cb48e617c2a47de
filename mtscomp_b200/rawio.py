# -*- coding: utf-8 -*-
"""Raw flat-binary loading (reference mtscomp.py:115-140)."""

import os.path as op
from pathlib import Path

import numpy as np


def load_raw_data(path=None, n_channels=None, dtype=None, offset=None, mmap=True):
    """Memory-map (or read) a flat binary file as an (n_samples, n_channels) array.

    Same contract as the reference: the file size must be a whole number of rows (ValueError otherwise), an empty
    file yields a (0, n_channels) array, `offset` is a byte offset of the data inside the file."""
    path = Path(path)
    assert path.exists(), "File %s does not exist." % path
    assert dtype, "The data type must be provided."
    n_channels = n_channels or 1
    offset = offset or 0
    itemsize = np.dtype(dtype).itemsize
    payload = op.getsize(str(path)) - offset
    n_samples = payload // (itemsize * n_channels)
    if n_samples * n_channels * itemsize != payload:
        raise ValueError(
            "The file size (%d bytes) is incompatible with the specified parameters "
            "(n_channels=%d, dtype=%s, offset=%d)" % (payload + offset, n_channels, dtype, offset))
    if n_samples * n_channels == 0:
        return np.zeros((0, n_channels), dtype=dtype)
    if mmap:
        return np.memmap(str(path), dtype=dtype, shape=(n_samples, n_channels), offset=offset)
    if offset > 0:  # pragma: no cover
        raise NotImplementedError()
    return np.fromfile(str(path), dtype).reshape((n_samples, n_channels))
