# -*- coding: utf-8 -*-
"""Extra measurement legs of bench.py for the BASELINE configs that are not the headline line:
configs[2] (one recording compressed by N ranks into ONE .cbin/.ch, verified), configs[3] (LFP band with spatial diff,
1 s chunks and a 0.1 s-chunk stress) and configs[4] (random-access latency of Reader slicing, p50/p99).
Every leg checks what it times against the source data / the oracle outside the timed regions."""
import hashlib
import json
import os
import shutil
import tempfile
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np


def scratch_dir(tag):
    base = '/dev/shm' if os.path.isdir('/dev/shm') and os.access('/dev/shm', os.W_OK) else tempfile.gettempdir()
    d = Path(base) / ('mtsb_%s_%d' % (tag, os.getpid()))
    d.mkdir(parents=True, exist_ok=True)
    return d


def pct(v, q):
    return float(np.percentile(np.asarray(v, dtype=np.float64), q))


# ---------------------------------------------------------------------------------------------------- configs[3]: LFP
def lfp_leg(cd, n_chunks, ns, threads, reps=3):
    """Device-resident compress / decompress of `n_chunks` LFP chunks of `ns` x 385 int16 with time + spatial diff."""
    import torch
    from mtscomp_b200 import _native, synth
    from oracle import codec as ora
    nc = 385
    fl = _native.flags_of(True, True, 'F')
    n_distinct = 8
    base = [np.ascontiguousarray(synth.lfp_chunk(ns, nc, seed=50 + i)) for i in range(n_distinct)]
    with ThreadPoolExecutor(threads) as ex:
        zs = list(ex.map(lambda x: ora.encode_chunk(x, True, True, 'F'), base))
    cb = ns * nc * 2
    raw = np.concatenate([base[i % n_distinct] for i in range(n_chunks)])
    rows = np.arange(n_chunks + 1, dtype=np.int64) * ns
    d_raw = torch.from_numpy(raw.view(np.uint8).reshape(-1)).cuda()
    cap = n_chunks * cd.compress_bound(ns, nc, 2, fl)
    d_comp = torch.empty(cap, dtype=torch.uint8, device='cuda')
    d_out = torch.empty(n_chunks * cb, dtype=torch.uint8, device='cuda')
    ref_offs = np.zeros(n_chunks + 1, dtype=np.int64)
    ref_offs[1:] = np.cumsum([len(zs[i % n_distinct]) for i in range(n_chunks)])
    d_ref = torch.from_numpy(np.frombuffer(b''.join(zs[i % n_distinct] for i in range(n_chunks)), np.uint8).copy()).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(fn):
        best, r = 1e30, None
        for _ in range(reps):
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        return best, r
    total = n_chunks * cb
    t_c, offs = timed(lambda: cd.compress_ptr(d_raw.data_ptr(), 1, rows, nc, 2, fl, d_comp.data_ptr(), 1, cap))
    t_g, st = timed(lambda: cd.decompress_ptr(d_comp.data_ptr(), 1, offs, rows, nc, 2, fl, d_out.data_ptr(), 1))
    ok_g = bool(torch.equal(d_out, d_raw)) and not st.any()
    d_out.zero_()
    t_r, st = timed(lambda: cd.decompress_ptr(d_ref.data_ptr(), 1, ref_offs, rows, nc, 2, fl, d_out.data_ptr(), 1))
    ok_r = bool(torch.equal(d_out, d_raw)) and not st.any()
    # the reference's decoder accepts the GPU streams of the first chunks
    comp = d_comp[:int(offs[-1])].cpu().numpy()
    for i in range(min(4, n_chunks)):
        assert zlib.decompress(bytes(comp[offs[i]:offs[i + 1]])) == ora.transform_chunk(base[i % n_distinct], True, True, 'F')
    assert ok_g and ok_r, 'LFP decode differs from the input'
    return {'chunks': n_chunks, 'chunk_samples': ns, 'chunk_bytes': cb, 'raw_GB': total / 1e9,
            'compress_GBps': total / t_c / 1e9, 'decompress_gpu_written_GBps': total / t_g / 1e9,
            'decompress_reference_written_GBps': total / t_r / 1e9,
            'gpu_size_over_zlib': int(offs[-1]) / int(ref_offs[-1]), 'timing': 'best of %d, host clock around synchronised calls' % reps,
            'exact': True}


# ---------------------------------------------------------------------------------------------------- configs[4]: latency
def write_reference_style(arr, sample_rate, cbin, ch, threads):
    """A file as the reference Writer writes it (oracle port: zlib level 6 per chunk, same .ch keys)."""
    from oracle import codec as ora
    blob, bounds, offsets = ora.encode_array(arr, sample_rate, 1.0, n_threads=threads)
    Path(cbin).write_bytes(blob)
    meta = {'version': '1.0', 'algorithm': 'zlib', 'comp_level': -1, 'do_time_diff': True, 'do_spatial_diff': False,
            'dtype': str(arr.dtype), 'n_channels': int(arr.shape[1]), 'sample_rate': float(sample_rate),
            'chunk_bounds': bounds, 'chunk_offsets': offsets, 'chunk_order': 'F',
            'sha1_compressed': hashlib.sha1(blob).hexdigest(), 'sha1_uncompressed': hashlib.sha1(arr.tobytes()).hexdigest(),
            'shape': list(arr.shape)}
    Path(ch).write_text(json.dumps(meta, indent=2, sort_keys=True))


def latency_leg(threads, n_chunks=16, n_gpu=(40, 40, 30, 12), n_cpu=(10, 10, 6, 3), seed=11, ns=30000, nc=384):
    """p50 / p99 of `r[t0:t1, :]` for 10 ms, 100 ms, 1 s and 10 s windows at random t0 on 384-channel 30 kHz data
    (23.04 MB chunks), cache_size = 1: the GPU Reader on a reference-written and on a GPU-written file, and the oracle
    port of the reference Reader on the reference-written file, all on this box, files on tmpfs."""
    import mtscomp_b200 as M
    from mtscomp_b200 import synth
    from oracle.reader import PortReader
    sr = float(ns)
    d = scratch_dir('lat')
    try:
        M.CONFIG_PATH = d / '.mtscomp'
        base = [synth.ap_chunk(ns, nc, seed=300 + i) for i in range(4)]
        arr = np.concatenate([base[i % 4] for i in range(n_chunks)])
        arr.tofile(d / 'np2.bin')
        write_reference_style(arr, sr, d / 'ref.cbin', d / 'ref.ch', threads)
        M.compress(d / 'np2.bin', d / 'gpu.cbin', d / 'gpu.ch', sample_rate=sr, n_channels=nc, dtype=np.int16,
                   quiet=True, check_after_compress=False)
        windows = [('10ms', ns // 100), ('100ms', ns // 10), ('1s', ns), ('10s', 10 * ns)]
        readers = [('gpu_reader_reference_written', lambda: M.decompress(d / 'ref.cbin', d / 'ref.ch', cache_size=1), n_gpu),
                   ('gpu_reader_gpu_written', lambda: M.decompress(d / 'gpu.cbin', d / 'gpu.ch', cache_size=1), n_gpu),
                   ('cpu_port_reader_reference_written', lambda: PortReader(d / 'ref.cbin', d / 'ref.ch', cache_size=1), n_cpu)]
        out = {'shape': '%d chunks of %d x %d int16, cache_size 1, files on %s' % (n_chunks, ns, nc, d.parent),
               'unit': 'ms', 'windows': {}}
        for name, mk, counts in readers:
            r = mk()
            r[0:300]                                            # first touch: context, staging buffers
            rng = np.random.default_rng(seed)
            for (wname, w), n in zip(windows, counts):
                ts = []
                for _ in range(n):
                    t0 = int(rng.integers(0, arr.shape[0] - w))
                    t = time.perf_counter()
                    a = r[t0:t0 + w, :]
                    ts.append((time.perf_counter() - t) * 1e3)
                    assert a.shape == (w, nc) and np.array_equal(a[::97], arr[t0:t0 + w:97])
                out['windows'].setdefault(wname, {})[name] = {'p50': pct(ts, 50), 'p99': pct(ts, 99), 'n': n}
            r.close()
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------- file to file through the API
def file_leg(n_chunks=128):
    """Writer.write and Reader.tofile through the drop-in API, files on tmpfs, against the format's own ceiling: the .ch
    records two SHA-1 digests, sequential passes at hashlib speed (measured here on the same bytes)."""
    import mtscomp_b200 as M
    from mtscomp_b200 import synth
    ns, nc, sr = 30000, 385, 30000.
    d = scratch_dir('file')
    try:
        M.CONFIG_PATH = d / '.mtscomp'
        base = [synth.ap_chunk(ns, nc, seed=700 + i) for i in range(4)]
        with open(d / 'rec.bin', 'wb') as f:
            for i in range(n_chunks):
                f.write(base[i % 4].tobytes())
        raw_bytes = n_chunks * ns * nc * 2
        # warm-up on a file of two batches (context, pinned staging buffers of the final size), then the timed runs
        with open(d / 'w.bin', 'wb') as f:
            for i in range(24):
                f.write(base[i % 4].tobytes())
        M.compress(d / 'w.bin', d / 'w.cbin', d / 'w.ch', sample_rate=sr, n_channels=nc, dtype=np.int16, quiet=True, check_after_compress=False)
        M.decompress(d / 'w.cbin', d / 'w.ch', d / 'w_back.bin', quiet=True, check_after_decompress=False).close()
        for p in ('w.bin', 'w.cbin', 'w.ch', 'w_back.bin'):
            (d / p).unlink()
        t = time.perf_counter()
        M.compress(d / 'rec.bin', d / 'rec.cbin', d / 'rec.ch', sample_rate=sr, n_channels=nc, dtype=np.int16, quiet=True,
                   check_after_compress=False)
        t_w = time.perf_counter() - t
        t = time.perf_counter()
        M.decompress(d / 'rec.cbin', d / 'rec.ch', d / 'back.bin', quiet=True, check_after_decompress=False).close()
        t_r = time.perf_counter() - t
        mm = np.memmap(d / 'rec.bin', dtype=np.uint8, mode='r')
        t = time.perf_counter()
        h = hashlib.sha1()
        for o in range(0, raw_bytes, 64 << 20):
            h.update(mm[o:o + (64 << 20)])
        t_h = time.perf_counter() - t
        meta = json.loads((d / 'rec.ch').read_text())
        assert meta['sha1_uncompressed'] == h.hexdigest()
        assert (d / 'back.bin').read_bytes() == bytes(mm)
        del mm
        return {'chunks': n_chunks, 'raw_GB': raw_bytes / 1e9, 'writer_write_GBps': raw_bytes / t_w / 1e9,
                'hashlib_sha1_GBps': raw_bytes / t_h / 1e9, 'writer_fraction_of_sha1': t_h / t_w,
                'reader_tofile_GBps': raw_bytes / t_r / 1e9, 'files': 'tmpfs' if str(d).startswith('/dev/shm') else 'tmp',
                'note': 'compress() / decompress(out=...) of the package, checks off; sha1_uncompressed verified, output file == input'}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------- configs[2]: one file, N ranks
def sharded_leg(rank, world, chunks_per_rank, group=None, zlib_sample=4, warm_chunks_per_rank=None):
    """A synthetic AP recording of world x chunks_per_rank one-second chunks on tmpfs, compressed by all ranks into ONE
    .cbin/.ch (sharding.write_sharded), then verified: every rank decodes its chunk range with the GPU Reader and
    compares it with the source; rank 0 inflates a sample of chunks with CPython zlib against the oracle's transform.
    Rank 0 returns the report, the others None."""
    import torch
    import torch.distributed as dist
    import mtscomp_b200 as M
    from mtscomp_b200 import sharding, synth
    from oracle import codec as ora
    ns, nc, sr = 30000, 385, 30000.
    n_chunks = world * chunks_per_rank
    tag = [None]
    if rank == 0:
        tag[0] = str(scratch_dir('shard'))
    if world > 1:
        dist.broadcast_object_list(tag, src=0, group=group)
    d = Path(tag[0])
    M.CONFIG_PATH = d / '.mtscomp'
    raw_path = d / 'rec.bin'
    cb = ns * nc * 2
    if rank == 0:
        with open(raw_path, 'wb') as f:
            f.truncate(n_chunks * cb)
    sharding._barrier(group)
    first, last = sharding.shard_range(n_chunks, rank, world)
    base = [synth.ap_chunk(ns, nc, seed=900 + i) for i in range(4)]
    fd = os.open(raw_path, os.O_WRONLY)
    for i in range(first, last):
        os.pwrite(fd, base[i % 4].tobytes(), i * cb)
    os.close(fd)
    sharding._barrier(group)
    # (first call: context and pinned staging buffers; the second one is timed.  For a large recording the warm-up
    # runs on a short recording of its own.)
    if warm_chunks_per_rank is None or warm_chunks_per_rank >= chunks_per_rank:
        sharding.write_sharded(raw_path, d / 'warm.cbin', d / 'warm.ch', rank, world, sample_rate=sr, n_channels=nc,
                               dtype=np.int16, group=group, hash_raw=False)
    else:
        wn = world * warm_chunks_per_rank
        if rank == 0:
            with open(d / 'warm.bin', 'wb') as f:
                for i in range(wn):
                    f.write(base[i % 4].tobytes())
        sharding._barrier(group)
        sharding.write_sharded(d / 'warm.bin', d / 'warm.cbin', d / 'warm.ch', rank, world, sample_rate=sr, n_channels=nc,
                               dtype=np.int16, group=group, hash_raw=False)
        sharding._barrier(group)
        if rank == 0:
            for name in ('warm.bin', 'warm.cbin', 'warm.ch'):
                os.unlink(d / name)
    sharding._barrier(group)
    t = time.perf_counter()
    offsets, secs = sharding.write_sharded(raw_path, d / 'rec.cbin', d / 'rec.ch', rank, world, sample_rate=sr,
                                           n_channels=nc, dtype=np.int16, group=group, hash_raw=True)
    wall = time.perf_counter() - t
    # verification, outside the timed part
    r = M.decompress(d / 'rec.cbin', d / 'rec.ch')
    src = np.memmap(raw_path, dtype=np.int16, mode='r').reshape(-1, nc)
    ok = True
    t = time.perf_counter()
    for lo in range(first, last, 16):
        hi = min(lo + 16, last)
        ok = ok and np.array_equal(r[lo * ns:hi * ns], src[lo * ns:hi * ns])
    dec_secs = time.perf_counter() - t
    r.close()
    flag = torch.tensor([1 if ok else 0, int(secs * 1e6), int(dec_secs * 1e6)], dtype=torch.int64, device='cuda')
    if world > 1:
        parts = [torch.zeros_like(flag) for _ in range(world)]
        dist.all_gather(parts, flag, group=group)
    else:
        parts = [flag]
    report = None
    if rank == 0:
        meta = json.loads((d / 'rec.ch').read_text())
        blob_fd = os.open(d / 'rec.cbin', os.O_RDONLY)
        sample = sorted(set(int(x) for x in np.linspace(0, n_chunks - 1, zlib_sample)))
        for i in sample:
            buf = os.pread(blob_fd, offsets[i + 1] - offsets[i], offsets[i])
            assert zlib.decompress(buf) == ora.transform_chunk(src[i * ns:(i + 1) * ns])
        os.close(blob_fd)
        all_ok = all(int(p[0]) == 1 for p in parts)
        assert all_ok, 'sharded file does not decode to the source'
        comp_secs = max(int(p[1]) for p in parts) / 1e6
        report = {'chunks': n_chunks, 'raw_GB': n_chunks * cb / 1e9, 'ranks': world,
                  'compressed_over_raw': offsets[-1] / (n_chunks * cb),
                  'compress_seconds_max_over_ranks': comp_secs,
                  'compress_GBps_memmap_to_host_parts': n_chunks * cb / comp_secs / 1e9,
                  'wall_seconds_with_pwrite_and_both_sha1': wall,
                  'gpu_reader_decode_GBps_to_host_arrays': n_chunks * cb / (max(int(p[2]) for p in parts) / 1e6) / 1e9,
                  'verified': 'every chunk decoded by the GPU Reader == source; chunks %s inflated by CPython zlib == oracle '
                              'transform; sha1_compressed / sha1_uncompressed recorded in the .ch' % sample,
                  'sha1_compressed': meta['sha1_compressed'], 'file': 'tmpfs'}
    del src
    sharding._barrier(group)
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
    return report


# ---------------------------------------------------------------------------------------------------- host link ceiling
def copy_ceiling(nbytes, barrier, reps=3):
    """Copy-only H2D and D2H of `nbytes` from/to pinned memory on this rank (all ranks at once): seconds (best of reps)."""
    import torch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dv = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    res = []
    for direction in (0, 1):
        best = 1e30
        for _ in range(reps):
            barrier()
            t = time.perf_counter()
            if direction == 0:
                dv.copy_(h, non_blocking=True)
            else:
                h.copy_(dv, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        res.append(best)
    del h, dv
    return res
