"""GPU: parity of the sm_100a kernels against the oracle and the reference-written golden files, through the C ABI.

Bit-exact everywhere (integer / byte work): K1 output == the bytes the reference hands to zlib; zlib (the reference
Reader's decoder) accepts every GPU-written chunk and returns those bytes; the GPU decoder returns exactly what the
reference Reader returns for reference-written chunks.  Compression ratio: GPU bytes <= 1.031 x zlib level 6 (north
star: ratio >= 0.97 x the reference's)."""
import hashlib
import json
import zlib

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import codec as ora

pytestmark = pytest.mark.gpu

MANIFEST = json.loads((GOLDEN / 'manifest.json').read_text())
CASES = sorted(MANIFEST['cases'])


def F(td=True, sd=False, order='F'):
    from mtscomp_b200 import _native
    return _native.flags_of(td, sd, order)


def load_case(name):
    m = MANIFEST['cases'][name]
    ch = json.loads((GOLDEN / (name + '.ch')).read_text())
    raw = np.fromfile(GOLDEN / (name + '.bin'), dtype=m['dtype']).reshape(m['shape'])
    return m, ch, raw, (GOLDEN / (name + '.cbin')).read_bytes()


def kw_of(ch):
    return dict(do_time_diff=ch['do_time_diff'], do_spatial_diff=ch['do_spatial_diff'], chunk_order=ch['chunk_order'])


def test_native_library_is_the_cuda_build(codec):
    from mtscomp_b200 import _native
    assert str(_native.LIB_PATH).endswith('libmtscomp_b200.so')
    assert codec.lib.mtsb_device_count() >= 1
    assert codec.get_param('sm_count') > 100


@pytest.mark.parametrize('name', CASES)
def test_golden_reference_files(codec, name):
    m, ch, raw, cbin = load_case(name)
    fl = F(**{'td': ch['do_time_diff'], 'sd': ch['do_spatial_diff'], 'order': ch['chunk_order']})
    b, o = ch['chunk_bounds'], ch['chunk_offsets']
    # K1 == reference transform bytes
    assert codec.delta_transform(raw[b[0]:b[1]], fl).tobytes() == (GOLDEN / (name + '.tr')).read_bytes()
    # K3+K4 on the reference-written .cbin == reference Reader output
    out, st = codec.decompress(cbin, o, b, ch['n_channels'], ch['dtype'], fl)
    assert not st.any()
    assert hashlib.sha1(out.tobytes()).hexdigest() == m['sha1_decoded']
    # K1+K2: zlib accepts each chunk and yields the reference's transform bytes
    comp, offs = codec.compress(raw, b, fl)
    for i in range(len(b) - 1):
        assert zlib.decompress(bytes(comp[offs[i]:offs[i + 1]])) == ora.transform_chunk(raw[b[i]:b[i + 1]], **kw_of(ch))
    out2, st2 = codec.decompress(comp, offs, b, ch['n_channels'], ch['dtype'], fl)
    assert not st2.any() and np.array_equal(out2, raw)


@pytest.mark.parametrize('dtype', ['uint8', 'int8', 'uint16', 'int16', 'int32', 'uint32', 'int64'])
@pytest.mark.parametrize('shape', [(1, 1), (7, 3), (64, 19), (65, 385), (257, 384), (130, 33), (3, 1000)])
def test_transform_all_flag_combinations(codec, dtype, shape):
    rng = np.random.default_rng(hash((dtype, shape)) % 2 ** 32)
    info = np.iinfo(dtype)
    x = rng.integers(info.min, info.max, shape, dtype=dtype, endpoint=True)
    for td in (True, False):
        for sd in (True, False):
            for order in 'FC':
                want = ora.transform_chunk(x, td, sd, order)
                assert codec.delta_transform(x, F(td, sd, order)).tobytes() == want
                back, ad = codec.inverse_transform(want, shape[0], shape[1], dtype, F(td, sd, order), want_adler=True)
                assert np.array_equal(back, x)
                assert ad == zlib.adler32(want)


@pytest.mark.parametrize('td,sd,order', [(True, False, 'F'), (True, True, 'F'), (True, False, 'C'), (False, False, 'F')])
def test_codec_ap_small_chunks(codec, td, sd, order):
    from mtscomp_b200 import synth
    x = synth.ap_chunk(ns=4000, nc=96, seed=31)
    rows = [0, 1500, 3000, 4000]
    fl = F(td, sd, order)
    comp, offs = codec.compress(x, rows, fl)
    for i in range(3):
        want = ora.transform_chunk(x[rows[i]:rows[i + 1]], td, sd, order)
        assert zlib.decompress(bytes(comp[offs[i]:offs[i + 1]])) == want
    out, st = codec.decompress(comp, offs, rows, 96, np.int16, fl)
    assert not st.any() and np.array_equal(out, x)
    # same chunks written by zlib (what the reference Writer produces) through the GPU decoder
    parts = [ora.encode_chunk(x[rows[i]:rows[i + 1]], td, sd, order) for i in range(3)]
    roffs = np.concatenate(([0], np.cumsum([len(p) for p in parts])))
    out2, st2 = codec.decompress(b''.join(parts), roffs, rows, 96, np.int16, fl)
    assert not st2.any() and np.array_equal(out2, x)


def test_full_size_ap_chunk_ratio_and_cross_compat(codec):
    """BASELINE config 1 chunk shape (30000 x 385): exact both ways + the north star's ratio bound."""
    from mtscomp_b200 import synth
    x = synth.ap_chunk(ns=30000, nc=385, seed=1234)
    want = ora.transform_chunk(x)
    comp, offs = codec.compress(x, [0, 30000], F())
    assert zlib.decompress(bytes(comp)) == want
    ref = zlib.compress(want)
    assert len(comp) <= 1.031 * len(ref), (len(comp), len(ref))
    out, st = codec.decompress(ref, [0, len(ref)], [0, 30000], 385, np.int16, F())
    assert not st.any() and np.array_equal(out, x)
    out2, st2 = codec.decompress(comp, offs, [0, 30000], 385, np.int16, F())
    assert not st2.any() and np.array_equal(out2, x)


def test_lfp_spatial_ratio(codec):
    from mtscomp_b200 import synth
    x = np.concatenate([synth.lfp_chunk(ns=2500, nc=385, seed=50 + i) for i in range(4)])
    rows = [0, 2500, 5000, 7500, 10000]
    comp, offs = codec.compress(x, rows, F(True, True))
    ref = sum(len(ora.encode_chunk(x[rows[i]:rows[i + 1]], True, True)) for i in range(4))
    assert len(comp) <= 1.031 * ref, (len(comp), ref)
    out, st = codec.decompress(comp, offs, rows, 385, np.int16, F(True, True))
    assert not st.any() and np.array_equal(out, x)


def test_many_chunks_batch_properties(codec):
    """Size-independent properties at a multi-sub-batch size: offsets monotone, every chunk independently decodable
    in any order (chunks are independent streams), deterministic bytes (reference tests.py:489-492 relies on it)."""
    from mtscomp_b200 import synth
    base = [synth.ap_chunk(ns=3000, nc=385, seed=70 + i) for i in range(3)]
    n = 40
    x = np.concatenate([base[i % 3] for i in range(n)])
    rows = np.arange(n + 1) * 3000
    codec.set_param('batch_bytes', 16 << 20)      # force several internal sub-batches
    try:
        comp, offs = codec.compress(x, rows, F())
        comp_b, offs_b = codec.compress(x, rows, F())
    finally:
        codec.set_param('batch_bytes', 2 << 30)
    assert np.array_equal(offs, offs_b) and np.array_equal(comp, comp_b)
    assert (np.diff(offs) > 0).all()
    # compressing the first 7 chunks alone gives the same bytes as the first 7 chunks of the big batch (chop-ability)
    comp7, offs7 = codec.compress(x[:7 * 3000], rows[:8], F())
    assert bytes(comp7) == bytes(comp[:offs[7]])
    # decode a permuted subset
    ids = [17, 3, 39, 0, 22]
    sub = b''.join(bytes(comp[offs[i]:offs[i + 1]]) for i in ids)
    soffs = np.concatenate(([0], np.cumsum([offs[i + 1] - offs[i] for i in ids])))
    out, st = codec.decompress(sub, soffs, np.arange(len(ids) + 1) * 3000, 385, np.int16, F())
    assert not st.any()
    for k, i in enumerate(ids):
        assert np.array_equal(out[k * 3000:(k + 1) * 3000], x[i * 3000:(i + 1) * 3000])


def test_inverse_single_pass_ragged_batches(codec):
    """K4's single-pass kernel: long look-back chains over many chunks at once, chunks of very different lengths in one
    launch, ragged last tiles, time and spatial sums, 2- and 8-byte elements, more channels than threads, repeated
    launches over the same cells."""
    from mtscomp_b200 import synth
    rng = np.random.default_rng(40)
    lens = [30000 - 7 * i for i in range(6)] + [50, 1]
    x = np.concatenate([synth.ap_chunk(ns=n, nc=385, seed=90 + i) for i, n in enumerate(lens)])
    rows = np.concatenate(([0], np.cumsum(lens)))
    for td, sd in ((True, False), (True, True)):
        comp, offs = codec.compress(x, rows, F(td, sd))
        for rep in range(2):
            out, st = codec.decompress(comp, offs, rows, 385, np.int16, F(td, sd))
            assert not st.any() and np.array_equal(out, x), (td, sd)
    y = np.cumsum(rng.integers(-9, 10, (5000, 100)), axis=0).astype(np.int64)
    comp, offs = codec.compress(y, [0, 3333, 5000], F())
    out, st = codec.decompress(comp, offs, [0, 3333, 5000], 100, np.int64, F())
    assert not st.any() and np.array_equal(out, y)
    z = np.cumsum(rng.integers(-9, 10, (4000, 1500)), axis=0).astype(np.int16)
    comp, offs = codec.compress(z, [0, 1000, 4000], F())
    out, st = codec.decompress(comp, offs, [0, 1000, 4000], 1500, np.int16, F())
    assert not st.any() and np.array_equal(out, z)


def test_incompressible_and_runs(codec):
    rng = np.random.default_rng(9)
    x = rng.integers(-32768, 32767, (5000, 64), dtype=np.int64).astype(np.int16)
    comp, offs = codec.compress(x, [0, 5000], F())
    assert len(comp) <= codec.compress_bound(5000, 64, 2, F())
    assert zlib.decompress(bytes(comp)) == ora.transform_chunk(x)
    z = np.zeros((30000, 8), np.int16)
    z[:, 7] = (np.arange(30000) // 15000) * 64
    comp, offs = codec.compress(z, [0, 30000], F())
    assert zlib.decompress(bytes(comp)) == ora.transform_chunk(z)
    assert len(comp) < 4000
    out, st = codec.decompress(comp, offs, [0, 30000], 8, np.int16, F())
    assert not st.any() and np.array_equal(out, z)


def test_corruption_is_detected_per_chunk(codec):
    _, ch, raw, cbin = load_case('nodiff')
    o = ch['chunk_offsets']
    bad = bytearray(cbin)
    bad[o[2] - 1] ^= 0x40       # adler32 trailer of chunk 1
    bad[o[2] + 40] ^= 0x10      # payload of chunk 2
    fl = F(ch['do_time_diff'], ch['do_spatial_diff'], ch['chunk_order'])
    _, st = codec.decompress(bytes(bad), o, ch['chunk_bounds'], ch['n_channels'], ch['dtype'], fl)
    assert st[0] == 0 and st[1] == 9 and st[2] != 0
    with pytest.raises(zlib.error):
        zlib.decompress(bytes(bad[o[1]:o[2]]))      # the reference rejects the same chunk
    # truncated stream and garbage
    _, st = codec.decompress(bytes(cbin[:o[1] - 9]), [0, o[1] - 9], ch['chunk_bounds'][:2], ch['n_channels'], ch['dtype'], fl)
    assert st[0] != 0
    junk = np.random.default_rng(1).integers(0, 256, 500, dtype=np.uint8).tobytes()
    _, st = codec.decompress(junk, [0, 500], ch['chunk_bounds'][:2], ch['n_channels'], ch['dtype'], fl)
    assert st[0] != 0


def test_zlib_variants_accepted(codec):
    """stored blocks, fixed-Huffman blocks, smaller declared windows, trailing bytes, multi-block streams."""
    from mtscomp_b200 import synth
    x = synth.ap_chunk(ns=2000, nc=32, seed=77)
    want = ora.transform_chunk(x)
    variants = []
    for level, wbits, strategy in [(0, 15, 0), (1, 15, 0), (9, 15, 0), (6, 9, 0), (6, 12, 0), (6, 15, zlib.Z_FIXED),
                                   (6, 15, zlib.Z_HUFFMAN_ONLY), (6, 15, zlib.Z_RLE)]:
        c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
        variants.append(c.compress(want) + c.flush())
    c = zlib.compressobj(6)
    parts = b''.join(c.compress(want[i:i + 9000]) + c.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(want), 9000))
    variants.append(parts + c.flush())
    variants.append(zlib.compress(want) + b'\x00\x01\x02\x03')
    for v in variants:
        assert zlib.decompress(v) == want
        out, st = codec.decompress(v, [0, len(v)], [0, 2000], 32, np.int16, F())
        assert not st.any() and np.array_equal(out, x)


def test_block_parallel_inflate_matches_serial(codec):
    """Index-less zlib streams (what the reference Writer emits) go through the block-parallel decoder
    (inflate_par.cuh); it must agree with the serial warp decoder and with the oracle, and report what it did."""
    from mtscomp_b200 import synth
    x = np.concatenate([synth.ap_chunk(ns=30000, nc=48, seed=90 + i) for i in range(3)])
    rows = [0, 30000, 60000, 90000]
    parts = [ora.encode_chunk(x[rows[i]:rows[i + 1]]) for i in range(3)]
    offs = np.concatenate(([0], np.cumsum([len(p) for p in parts])))
    blob = b''.join(parts)
    try:
        codec.set_param('par_inflate', 1)
        codec.set_param('par_cells', 0)
        for wide in (0, 1):                      # both shapes of the chain-of-tiles resolve kernel
            codec.set_param('par_lz_wide', wide)
            outw, stw = codec.decompress(blob, offs, rows, 48, np.int16, F())
            assert not stw.any() and np.array_equal(outw, x) and codec.get_param('par_resumed') == 3
        codec.set_param('par_lz_wide', -1)
        codec.set_param('par_cells', 1)          # blocks resolved in parallel into cells (the few-streams path)
        outc, stc = codec.decompress(blob, offs, rows, 48, np.int16, F())
        assert not stc.any() and np.array_equal(outc, x) and codec.get_param('par_resumed') == 3
        codec.set_param('par_cells', -1)
        out1, st1 = codec.decompress(blob, offs, rows, 48, np.int16, F())
        chained, resumed = codec.get_param('par_chained'), codec.get_param('par_resumed')
        codec.set_param('par_inflate', 0)
        out0, st0 = codec.decompress(blob, offs, rows, 48, np.int16, F())
        assert codec.get_param('par_chained') == 0
    finally:
        codec.set_param('par_inflate', 1)
        codec.set_param('par_lz_wide', -1)
        codec.set_param('par_cells', -1)
    assert not st1.any() and not st0.any()
    assert np.array_equal(out1, x) and np.array_equal(out0, x)
    assert resumed == 3 and chained >= 3 * 30          # ~39 zlib blocks per 2.9 MB stream


def test_block_parallel_inflate_reports_corruption(codec):
    from mtscomp_b200 import synth
    x = synth.ap_chunk(ns=30000, nc=32, seed=95)
    good = ora.encode_chunk(x)
    for pos in (len(good) // 3, len(good) - 2, 1):
        bad = bytearray(good)
        bad[pos] ^= 0x5a
        with pytest.raises(zlib.error):
            zlib.decompress(bytes(bad))
        _, st = codec.decompress(bytes(bad), [0, len(bad)], [0, 30000], 32, np.int16, F())
        assert st[0] != 0, pos
    out, st = codec.decompress(good + b'\x00' * 7, [0, len(good) + 7], [0, 30000], 32, np.int16, F())
    assert st[0] == 0 and np.array_equal(out, x)        # trailing bytes are ignored, as zlib does


def test_block_parallel_skips_a_false_candidate(codec):
    """Chunk seed 101 of the bench workload has one bit pattern inside a block that passes the full header validation
    (314 candidates for 313 blocks).  The block before it must run on to its real end and the chain must skip it."""
    from mtscomp_b200 import synth
    x = np.ascontiguousarray(synth.ap_chunk(30000, 385, seed=101))
    good = ora.encode_chunk(x)
    out, st = codec.decompress(good, [0, len(good)], [0, 30000], 385, np.int16, F())
    assert st[0] == 0 and np.array_equal(out, x)
    cands, chained = codec.get_param('par_candidates'), codec.get_param('par_chained')
    assert codec.get_param('par_resumed') == 1 and chained >= 300
    assert cands >= chained                     # (cands > chained on this chunk; >= keeps the test data-independent)


def test_indexed_segments_block_kernels_match_serial(codec):
    """GPU-written chunks: the indexed segments decode through the second-format kernels (seg_tokens / seg_resolve, the
    default), through the block kernels (seg_v2 = 0) or through the serial warp decoder (par_indexed = 0 as well); all
    must give the input back, including chunks with segments stored uncompressed."""
    from mtscomp_b200 import synth
    rng = np.random.default_rng(12)
    x = np.concatenate([synth.ap_chunk(ns=20000, nc=64, seed=70),
                        rng.integers(-32768, 32767, (20000, 64)).astype(np.int16),      # incompressible: stored segments
                        synth.ap_chunk(ns=20000, nc=64, seed=71)])
    rows = [0, 20000, 40000, 60000]
    comp, offs = codec.compress(x, rows, F())
    try:
        codec.set_param('par_indexed', 0)
        out2, st2 = codec.decompress(comp, offs, rows, 64, np.int16, F())
        resumed2 = codec.get_param('par_resumed')
        codec.set_param('seg_v2', 0)
        codec.set_param('par_indexed', 1)
        out1, st1 = codec.decompress(comp, offs, rows, 64, np.int16, F())
        resumed = codec.get_param('par_resumed')
        codec.set_param('par_indexed', 0)
        out0, st0 = codec.decompress(comp, offs, rows, 64, np.int16, F())
        assert codec.get_param('par_resumed') == 0
    finally:
        codec.set_param('par_indexed', 1)
        codec.set_param('seg_v2', 1)
    assert not st1.any() and not st0.any() and not st2.any()
    assert np.array_equal(out1, x) and np.array_equal(out0, x) and np.array_equal(out2, x)
    assert resumed >= 10 and resumed2 >= 10     # the segments of the two compressible chunks went through the parallel kernels
    bad = bytearray(comp)
    bad[offs[0] + (offs[1] - offs[0]) // 2] ^= 0x11
    _, st = codec.decompress(bytes(bad), offs, rows, 64, np.int16, F())
    assert st[0] != 0 and st[1] == 0 and st[2] == 0


def test_block_parallel_long_codes(codec):
    """Heavy-tailed data gives Huffman codes longer than the fast tables (10 / 8 bits): the limit-word path."""
    rng = np.random.default_rng(5)
    for x in ((rng.laplace(0, 40, (60000, 6))).astype(np.int16),
              np.where(rng.random((60000, 6)) < 0.02, rng.integers(-30000, 30000, (60000, 6)),
                       rng.normal(0, 3, (60000, 6))).astype(np.int16)):
        good = ora.encode_chunk(x)
        out, st = codec.decompress(good, [0, len(good)], [0, 60000], 6, np.int16, F())
        assert st[0] == 0 and np.array_equal(out, x) and codec.get_param('par_resumed') == 1


def test_gpu_stream_bytes_equal_the_host_emulation(codec):
    """The encoder's output must not depend on how the hardware orders same-address shared-memory stores or schedules
    warps: the same kernel sources compiled for the host (csrc/emu, cooperative fibers, a completely different
    schedule) must produce byte-identical streams.  Runs/zeros make every unit of a batch collide in the hash tables."""
    from mtscomp_b200 import _native, build, synth
    try:
        emu = _native.Codec(0, lib=_native.load_library(build.build_emulation()))
    except Exception as e:                                   # no host compiler on the box: nothing to compare with
        pytest.skip('emulation build unavailable: %s' % e)
    z = np.zeros((2000, 16), dtype=np.int16); z[:, 3] = 17; z[1000:, 5] = -3
    rng = np.random.default_rng(3)
    flat = np.repeat(rng.integers(-200, 200, (40, 12)), 100, axis=0).astype(np.int16)       # plateaus
    for x, rows in ((z, [0, 1000, 2000]), (flat, [0, 4000]), (synth.ap_chunk(ns=3000, nc=24, seed=5), [0, 3000])):
        g, go = codec.compress(x, rows, F())
        e, eo = emu.compress(x, rows, F())
        assert list(go) == list(eo) and bytes(g) == bytes(e)


def test_block_parallel_marker_chains(codec):
    """Period-20 KB data: every deflate block consists of references into the block before it, so in the cells path
    (blocks resolved in parallel) a marker has to be chased through several blocks."""
    rng = np.random.default_rng(9)
    pat = rng.integers(-3000, 3000, 10000).astype(np.int16)
    x = np.tile(pat, 900).reshape(-1, 1)
    good = ora.encode_chunk(x)
    try:
        for cells in (1, 0):
            codec.set_param('par_cells', cells)
            out, st = codec.decompress(good, [0, len(good)], [0, x.shape[0]], 1, np.int16, F())
            assert st[0] == 0 and np.array_equal(out, x)
            assert codec.get_param('par_resumed') == 1 and codec.get_param('par_chained') >= 3
    finally:
        codec.set_param('par_cells', -1)


FLOAT_CASES = sorted(json.loads((GOLDEN / 'manifest.json').read_text()).get('float_cases', {}))


@pytest.mark.parametrize('name', FLOAT_CASES)
def test_float_golden_reference_files(codec, name):
    """float32 / float64: the CUDA path returns what the reference Reader returns, bit for bit (also where that
    differs from the input: float32_wild), for the reference-written file and for its own streams; the reference's
    decoder (zlib) gets the reference's transform bytes out of the GPU-written streams."""
    from mtscomp_b200 import _native
    m = json.loads((GOLDEN / 'manifest.json').read_text())['float_cases'][name]
    ch = json.loads((GOLDEN / (name + '.ch')).read_text())
    raw = np.fromfile(GOLDEN / (name + '.bin'), dtype=m['dtype']).reshape(m['shape'])
    dec = np.fromfile(GOLDEN / (name + '.dec'), dtype=m['dtype']).reshape(m['shape'])
    cbin = (GOLDEN / (name + '.cbin')).read_bytes()
    fl = _native.flags_of(ch['do_time_diff'], ch['do_spatial_diff'], ch['chunk_order'])
    b = ch['chunk_bounds']
    out, st = codec.decompress(cbin, ch['chunk_offsets'], b, ch['n_channels'], raw.dtype, fl)
    assert not st.any() and out.tobytes() == dec.tobytes()
    comp, offs = codec.compress(raw, b, fl)
    for i in range(len(b) - 1):
        assert zlib.decompress(bytes(comp[offs[i]:offs[i + 1]])) == ora.transform_chunk(
            raw[b[i]:b[i + 1]], ch['do_time_diff'], ch['do_spatial_diff'], ch['chunk_order'])
    out2, st2 = codec.decompress(comp, offs, b, ch['n_channels'], raw.dtype, fl)
    assert not st2.any() and out2.tobytes() == dec.tobytes()
    assert codec.delta_transform(raw[b[0]:b[1]], fl).tobytes() == (GOLDEN / (name + '.tr')).read_bytes()
