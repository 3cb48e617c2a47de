"""GPU: the cases VERDICT round 1 listed as untested -- the full codec at BASELINE configs[4]'s shape, the serial
fallback for streams of 256 MB and more, host-buffer decode across several sub-batches, a fuzz of valid zlib streams
(levels x strategies x window sizes; distance 32768, length 258), and an index that is self-consistent but false."""
import struct
import zlib

import numpy as np
import pytest

from oracle import codec as ora

pytestmark = pytest.mark.gpu


def F(td=True, sd=False, order='F'):
    from mtscomp_b200 import _native
    return _native.flags_of(td, sd, order)


def test_full_codec_at_np2_shape(codec):
    """30000 x 384 int16 (23.04 MB chunks): GPU-written chunks inflate with zlib to the oracle's transform bytes, stay
    within 3.1 % of zlib's size, and both kinds of stream decode exactly on the GPU."""
    from mtscomp_b200 import synth
    x = np.concatenate([synth.ap_chunk(30000, 384, seed=310 + i) for i in range(2)])
    rows = [0, 30000, 60000]
    comp, offs = codec.compress(x, rows, F())
    ref = [ora.encode_chunk(x[rows[i]:rows[i + 1]]) for i in range(2)]
    for i in range(2):
        assert zlib.decompress(bytes(comp[offs[i]:offs[i + 1]])) == ora.transform_chunk(x[rows[i]:rows[i + 1]])
    assert len(comp) <= 1.031 * sum(map(len, ref))
    out, st = codec.decompress(comp, offs, rows, 384, np.int16, F())
    assert not st.any() and np.array_equal(out, x)
    roffs = [0, len(ref[0]), len(ref[0]) + len(ref[1])]
    out, st = codec.decompress(b''.join(ref), roffs, rows, 384, np.int16, F())
    assert not st.any() and np.array_equal(out, x)
    assert codec.get_param('par_resumed') == 2


def test_streams_of_256mb_and_more_take_the_serial_decoder(codec):
    """A reference-style stream of more than 2^28 bytes (incompressible data, stored blocks) bypasses the block kernels
    (32-bit bit offsets) and must still decode exactly; so must the GPU-written form of the same chunk."""
    rng = np.random.default_rng(5)
    x = rng.integers(-32768, 32767, (70_000_000, 2), dtype=np.int16)             # 280 MB
    z = ora.encode_chunk(x)
    assert len(z) >= 1 << 28
    out, st = codec.decompress(z, [0, len(z)], [0, x.shape[0]], 2, np.int16, F())
    assert not st.any() and np.array_equal(out, x)
    assert codec.get_param('par_resumed') == 0
    comp, offs = codec.compress(x, [0, x.shape[0]], F())
    out, st = codec.decompress(comp, offs, [0, x.shape[0]], 2, np.int16, F())
    assert not st.any() and np.array_equal(out, x)


def test_host_buffer_decode_across_many_sub_batches(codec):
    """Host compressed bytes -> host array with sub-batches far smaller than the call (reference-written and GPU-written
    chunks mixed in one call): the pipelined copies must deliver every chunk to its place."""
    from mtscomp_b200 import synth
    ns, nc, n = 9000, 96, 12
    x = np.concatenate([synth.ap_chunk(ns, nc, seed=500 + i) for i in range(n)])
    rows = [i * ns for i in range(n + 1)]
    comp, offs = codec.compress(x, rows, F())
    parts = []
    for i in range(n):
        parts.append(bytes(comp[offs[i]:offs[i + 1]]) if i % 3 else ora.encode_chunk(x[rows[i]:rows[i + 1]]))
    moffs = np.concatenate(([0], np.cumsum([len(p) for p in parts])))
    try:
        codec.set_param('par_batch_bytes', 1 << 20)
        codec.set_param('host_batch_bytes', 1 << 20)
        codec.set_param('batch_bytes', 2 << 20)
        out, st = codec.decompress(b''.join(parts), moffs, rows, nc, np.int16, F())
    finally:
        codec.set_param('par_batch_bytes', 2 << 30)
        codec.set_param('host_batch_bytes', 512 << 20)
        codec.set_param('batch_bytes', 2 << 30)
    assert not st.any() and np.array_equal(out, x)


def _fuzz_inputs():
    rng = np.random.default_rng(99)
    n = 400_000
    yield 'random', rng.integers(0, 256, n, dtype=np.uint8)
    yield 'lowentropy', rng.integers(0, 4, n, dtype=np.uint8)
    yield 'periodic7', np.tile(np.arange(7, dtype=np.uint8), n // 7 + 1)[:n]
    yield 'runs', np.repeat(rng.integers(0, 256, n // 300 + 1, dtype=np.uint8), 300)[:n]            # length-258 matches
    far = rng.integers(0, 256, 32768, dtype=np.uint8)
    yield 'distance32768', np.concatenate([far, far, far[:1000], rng.integers(0, 256, 5000, dtype=np.uint8), far])
    yield 'text', np.frombuffer((b'the quick brown fox jumps over the lazy dog. ' * 9000)[:n], dtype=np.uint8).copy()


def test_fuzz_of_valid_zlib_streams(codec):
    """Whatever zlib can write, the GPU decoder must read: levels 0..9, every strategy, window sizes 9..15, memLevel 1..9
    (many small blocks), on six kinds of data -- each stream fed as one chunk of uint8 data without differences."""
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]
    streams, datas = [], []
    for name, data in _fuzz_inputs():
        for level, strat, wbits, mem in [(0, 0, 15, 8), (1, 0, 15, 8), (6, 0, 15, 8), (9, 0, 15, 9), (6, 1, 15, 8), (6, 2, 15, 8),
                                         (6, 3, 15, 8), (6, 4, 15, 8), (9, 0, 9, 1), (4, 0, 12, 3), (2, 3, 10, 5)]:
            co = zlib.compressobj(level, zlib.DEFLATED, wbits, mem, strategies[strat])
            z = co.compress(data.tobytes()) + co.flush()
            assert zlib.decompress(z) == data.tobytes()
            streams.append(z)
            datas.append(data)
    n = len(datas[0])
    same = [i for i in range(len(datas)) if len(datas[i]) == n]
    other = [i for i in range(len(datas)) if len(datas[i]) != n]
    for group in (same, other):
        # (equal-length chunks go in one batched call; the rest one by one)
        calls = [group] if group is same else [[i] for i in group]
        for ids in calls:
            blob = b''.join(streams[i] for i in ids)
            offs = np.concatenate(([0], np.cumsum([len(streams[i]) for i in ids])))
            rows = np.concatenate(([0], np.cumsum([len(datas[i]) for i in ids])))
            out, st = codec.decompress(blob, offs, rows, 1, np.uint8, 0)
            assert not st.any(), [ids[k] for k in np.flatnonzero(st)]
            assert np.array_equal(out[:, 0], np.concatenate([datas[i] for i in ids]))


def test_a_false_but_self_consistent_index_is_not_trusted(codec):
    """Bytes after a reference-written stream that pass every check of both index formats (magic, k, lengths tiling the
    stream, checksum word, table sizes): zlib ignores them, so the chunk is valid and must decode to the same data."""
    from mtscomp_b200 import synth
    ns, nc = 6000, 24
    x = synth.ap_chunk(ns, nc, seed=77)
    z = ora.encode_chunk(x)
    raw, seg = x.nbytes, 252000                                   # 21 channel runs of 12000 bytes
    k = (raw + seg - 1) // seg
    body = len(z) - 8
    lens = [body // k] * k
    lens[-1] += body - sum(lens)
    v1 = z + b''.join(struct.pack('<I', v) for v in lens) + struct.pack('<IIII', seg, k, 0x4253544D, sum(lens))
    n_sub = sum((min(seg, raw - j * seg) + 8191) // 8192 + (1 if min(seg, raw - j * seg) > 4096 else 0) for j in range(k))
    table = b''.join(struct.pack('<I', 20000 | (3 << 17)) for _ in range(n_sub))
    v2 = z + table + b''.join(struct.pack('<I', v) for v in lens) + struct.pack('<II', 8192, 1024) + \
        struct.pack('<IIII', seg, k, 0x3253544D, sum(lens))
    for forged in (v1, v2):
        assert zlib.decompress(forged) == ora.transform_chunk(x)
        out, st = codec.decompress(forged, [0, len(forged)], [0, ns], nc, np.int16, F())
        assert not st.any() and np.array_equal(out, x)
    # ... while a genuinely damaged stream under a genuine index is still reported
    comp, offs = codec.compress(x, [0, ns], F())
    bad = bytearray(comp)
    bad[len(bad) // 3] ^= 0x40
    _, st = codec.decompress(bytes(bad), offs, [0, ns], nc, np.int16, F())
    assert st[0] != 0
