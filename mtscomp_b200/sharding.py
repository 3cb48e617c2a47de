# -*- coding: utf-8 -*-
"""Chunk-range sharding across GPUs: contiguous ranges per rank, host-side gather of compressed sizes.

Chunks are independent zlib streams, so there is no collective on the data path.  The only cross-rank datum is the
per-chunk compressed size, gathered as Python objects over whatever process group exists (gloo on CPU, nccl ranks use
the default group's object collectives) and prefix-summed into the reference's `chunk_offsets` (mtscomp.py:453-480).
"""

import numpy as np


def shard_range(n_chunks, rank, world):
    """Chunks [first, last) of rank `rank`: contiguous, in file order, sizes differ by at most ceil/floor."""
    per = -(-n_chunks // world)
    return min(rank * per, n_chunks), min((rank + 1) * per, n_chunks)


def assemble_offsets(sizes_by_rank):
    """sizes_by_rank: list (rank order) of per-chunk compressed sizes -> global chunk_offsets (n_chunks + 1)."""
    sizes = [int(s) for part in sizes_by_rank for s in part]
    return [0] + np.cumsum(sizes, dtype=np.int64).tolist() if sizes else [0]


def rank_base_offsets(sizes_by_rank):
    """Byte offset in the .cbin at which each rank's packed output starts."""
    totals = [int(np.sum(p, dtype=np.int64)) if len(p) else 0 for p in sizes_by_rank]
    return [0] + np.cumsum(totals, dtype=np.int64).tolist()[:-1]


def gather_sizes(local_sizes, group=None):
    """All ranks get every rank's per-chunk sizes (torch.distributed object all-gather; works on gloo and nccl)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [list(map(int, local_sizes))]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(map(int, local_sizes)), group=group)
    return out


def compress_sharded(codec, data, chunk_bounds, flags, rank, world, group=None):
    """Compress this rank's chunk range of `data` (rows chunk_bounds[first]..chunk_bounds[last]) and return
    (packed bytes of this rank, global chunk_offsets, byte offset of this rank's part in the .cbin)."""
    first, last = shard_range(len(chunk_bounds) - 1, rank, world)
    if last > first:
        b0 = chunk_bounds[first]
        rows = np.asarray(chunk_bounds[first:last + 1], dtype=np.int64) - b0
        comp, offs = codec.compress(data[b0:chunk_bounds[last]], rows, flags)
        sizes = np.diff(offs).tolist()
    else:
        comp, sizes = np.zeros(0, np.uint8), []
    all_sizes = gather_sizes(sizes, group)
    return comp, assemble_offsets(all_sizes), rank_base_offsets(all_sizes)[rank] if world > 1 else 0
