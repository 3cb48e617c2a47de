#!/usr/bin/env python
"""bench.py — compress / decompress GB/s of raw int16 through the per-chunk codec on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE configs[1] shape — a 10-minute 385-channel 30 kHz AP recording, 600 one-second
chunks of 23.1 MB (13.86 GB raw) per GPU, synthetic (seeded band-limited noise + spikes, mtscomp_b200/synth.py), built
from `n_distinct` distinct chunks tiled in order.  One step = one pass of the codec over all the chunks:
  value            compress, chunks and output resident in HBM (CUDA events on the codec's stream)
  e2e              compress through the C ABI with pinned HOST buffers (H2D of the raw chunks and D2H of the .cbin bytes
                   inside the timed region)
  decompress.*     decode of reference-written streams (CPU zlib level 6 = what the reference Writer emits) and of
                   GPU-written streams, device-resident and end to end
Chunks shard across GPUs as independent ranges with no collective on the data path (weak scaling: every rank runs
the same 600-chunk workload); torch.distributed is only used for the barrier and the max-over-ranks of the times.

`--impl reference` times the reference's own CPU path for the same metric (oracle port: NumPy diff + zlib in a thread
pool, mtscomp.py:375-423 / 602-650) on a bounded sample with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NS, NC = 30000, 385
CHUNK_BYTES = NS * NC * 2
METRIC = 'compress_GBps_raw_int16'
UNIT = 'GB/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--chunks', type=int, default=600, help='chunks per GPU per step (600 = 10 min of AP data)')
    ap.add_argument('--distinct', type=int, default=8, help='distinct synthetic chunks tiled to --chunks')
    ap.add_argument('--cpu-sample', type=int, default=0, help='chunks in the CPU baseline sample (0 = auto)')
    ap.add_argument('--no-legs', action='store_true', help='skip the configs[2..4] legs (sharded file, LFP, latency)')
    ap.add_argument('--shard-chunks', type=int, default=24, help='configs[2] leg: chunks per rank of the one-file recording')
    ap.add_argument('--lfp-chunks', type=int, default=1200, help='configs[3] leg: 1 s LFP chunks (N=1 only)')
    ap.add_argument('--lfp-small-chunks', type=int, default=12000, help='configs[3] leg: 0.1 s LFP chunks (N=1 only)')
    ap.add_argument('--latency-chunks', type=int, default=16, help='configs[4] leg: chunks of the 384-channel files (N=1 only)')
    return ap.parse_args()


def workload_config(n_chunks, n_distinct):
    """The same dict for both arms (--impl b200 / reference): what is measured, not how."""
    return {'workload': 'BASELINE configs[1] shape: 10 min AP, 385ch x 30kHz int16, %d chunks of 1 s (23.1 MB) '
                        'per GPU, chunk_order F, time diff; compress AND decompress' % n_chunks,
            'chunks_per_gpu': n_chunks, 'n_distinct_chunks': n_distinct, 'raw_bytes_per_gpu': n_chunks * CHUNK_BYTES,
            'l2': 'inputs larger than L2 (%.0f MB distinct raw per GPU, %.1f GB per step)' % (
                n_distinct * CHUNK_BYTES / 1e6, n_chunks * CHUNK_BYTES / 1e9),
            'sharding': 'contiguous chunk ranges per GPU, no collective on the data path'}


def lz77_profile():
    """Figures of the committed ncu capture of the dominant kernel (profiles/r02_lz77_ncu.json)."""
    tp = ROOT / 'profiles' / 'r02_lz77_ncu.json'
    return json.loads(tp.read_text()) if tp.exists() else {}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# --------------------------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's threaded path
# --------------------------------------------------------------------------------------------------------------------

def cpu_codec_sample(base, n_sample, threads):
    """Reference-style batch: one chunk per thread through np.diff/tobytes('F')/zlib.compress, then the inverse."""
    from oracle import codec as ora
    chunks = [base[i % len(base)] for i in range(n_sample)]
    with ThreadPoolExecutor(threads) as ex:
        t = time.perf_counter()
        comp = list(ex.map(ora.encode_chunk, chunks))
        t_c = time.perf_counter() - t
        t = time.perf_counter()
        dec = list(ex.map(lambda c: ora.decode_chunk(c, NS, NC, np.int16), comp))
        t_d = time.perf_counter() - t
    assert np.array_equal(dec[0], chunks[0])
    raw = n_sample * CHUNK_BYTES
    return raw / t_c / 1e9, raw / t_d / 1e9, comp


def run_reference(args):
    from mtscomp_b200 import synth
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    threads = host_threads()
    n_sample = args.cpu_sample or max(2, min(threads, 32))
    base = [synth.ap_chunk(NS, NC, seed=1234 + i, t0=i * NS) for i in range(min(args.distinct, n_sample))]
    vals_c, vals_d = [], []
    for it in range(args.warmup + args.steps):
        c, d, _ = cpu_codec_sample(base, n_sample, threads)
        if it >= args.warmup:
            vals_c.append(c)
            vals_d.append(d)
    v = float(np.mean(vals_c))
    sample = '%d chunks (%d distinct) of %dx%d int16 per step, %d zlib threads, in memory' % (
        n_sample, len(base), NS, NC, threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': n_sample * CHUNK_BYTES / v / 1e6, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int16', 'data': 'synthetic',
        'config': workload_config(args.chunks, min(args.distinct, args.chunks)),
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample,
                         'decompress_value': float(np.mean(vals_d))},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'decompress': {'value': float(np.mean(vals_d)), 'unit': UNIT},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist
    from mtscomp_b200 import _native, synth

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor(x, dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    stream = torch.cuda.Stream()
    cd = _native.Codec(local, stream=stream.cuda_stream)
    lib = cd.lib
    fl = _native.TIME_DIFF
    n_chunks = args.chunks
    n_distinct = max(1, min(args.distinct, n_chunks))
    threads = host_threads()

    # ---- data: distinct seeded chunks, tiled; reference-written streams of the same chunks (CPU zlib)
    base = [synth.ap_chunk(NS, NC, seed=1234 + i + 1000 * rank, t0=i * NS) for i in range(n_distinct)]
    t0 = time.perf_counter()
    n_cpu = max(n_distinct, args.cpu_sample or min(threads, 32))   # every distinct chunk is needed as a reference stream
    cpu_c, cpu_d, ref_streams = cpu_codec_sample(base, n_cpu, threads) if rank == 0 else (None, None, None)
    cpu_secs = time.perf_counter() - t0
    if rank != 0:
        from oracle import codec as ora
        with ThreadPoolExecutor(threads) as ex:
            ref_streams = list(ex.map(ora.encode_chunk, base))
    ref_streams = ref_streams[:n_distinct]
    raw_bytes = n_chunks * CHUNK_BYTES
    rows = np.arange(n_chunks + 1, dtype=np.int64) * NS

    # the end-to-end legs need pinned host buffers (raw, .cbin bound, decoded, reference .cbin: ~78 MB per chunk); they
    # run on all the chunks unless this rank's share of the free host memory is smaller
    n_e2e = n_chunks
    try:
        import psutil
        avail = psutil.virtual_memory().available * 0.55 / max(world, 1)
        n_e2e = int(max(8, min(n_chunks, avail // (3.4 * CHUNK_BYTES))))
    except Exception:
        pass
    e2e_bytes = n_e2e * CHUNK_BYTES
    rows_e = rows[:n_e2e + 1]
    # device-resident copies (kernel path): tiled on the device from the distinct chunks
    cap = n_chunks * cd.compress_bound(NS, NC, 2, fl)
    base_dev = [torch.from_numpy(b.reshape(-1).view(np.uint8)).cuda() for b in base]
    d_raw = torch.empty(raw_bytes, dtype=torch.uint8, device='cuda')
    for i in range(n_chunks):
        d_raw[i * CHUNK_BYTES:(i + 1) * CHUNK_BYTES].copy_(base_dev[i % n_distinct])
    d_comp = torch.empty(cap, dtype=torch.uint8, device='cuda')
    d_out = torch.empty(raw_bytes, dtype=torch.uint8, device='cuda')
    ref_sizes = [len(ref_streams[i % n_distinct]) for i in range(n_chunks)]
    ref_offs = np.concatenate(([0], np.cumsum(ref_sizes))).astype(np.int64)
    ref_dev = [torch.from_numpy(np.frombuffer(r, np.uint8).copy()).cuda() for r in ref_streams]
    d_ref = torch.empty(int(ref_offs[-1]) + 64, dtype=torch.uint8, device='cuda')
    for i in range(n_chunks):
        d_ref[ref_offs[i]:ref_offs[i + 1]].copy_(ref_dev[i % n_distinct])
    # pinned host copies (end-to-end path) of the first n_e2e chunks
    cap_e = n_e2e * cd.compress_bound(NS, NC, 2, fl)
    ref_offs_e = ref_offs[:n_e2e + 1]
    h_raw_t = torch.empty(e2e_bytes, dtype=torch.uint8, pin_memory=True)
    h_raw = h_raw_t.numpy()
    for i in range(n_e2e):
        h_raw[i * CHUNK_BYTES:(i + 1) * CHUNK_BYTES] = base[i % n_distinct].reshape(-1).view(np.uint8)
    h_comp_t = torch.empty(cap_e, dtype=torch.uint8, pin_memory=True)
    h_out_t = torch.empty(e2e_bytes, dtype=torch.uint8, pin_memory=True)
    h_ref_t = torch.empty(int(ref_offs_e[-1]) + 64, dtype=torch.uint8, pin_memory=True)
    h_ref = h_ref_t.numpy()
    for i in range(n_e2e):
        h_ref[ref_offs[i]:ref_offs[i + 1]] = np.frombuffer(ref_streams[i % n_distinct], np.uint8)
    del base_dev, ref_dev
    torch.cuda.synchronize()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(fn):
        with torch.cuda.stream(stream):
            ev[0].record(stream)
            r = fn()
            ev[1].record(stream)
        ev[1].synchronize()
        return ev[0].elapsed_time(ev[1]), r

    state = {}

    def step(record):
        # (a) compress, device resident
        ms, offs = timed(lambda: cd.compress_ptr(d_raw.data_ptr(), 1, rows, NC, 2, fl, d_comp.data_ptr(), 1, cap))
        tm, ln = cd.timings(), cd.launches()
        state['offs'] = offs
        # (b) compress end to end from pinned host memory
        ms_e, offs_e = timed(lambda: cd.compress_ptr(h_raw_t.data_ptr(), 0, rows_e, NC, 2, fl, h_comp_t.data_ptr(), 0, cap_e))
        ln += cd.launches()
        # (c) decompress reference-written streams, device resident
        ms_r, st = timed(lambda: cd.decompress_ptr(d_ref.data_ptr(), 1, ref_offs, rows, NC, 2, fl, d_out.data_ptr(), 1))
        tm_r = cd.timings()
        ln += cd.launches()
        assert not st.any()
        # (d) decompress GPU-written streams, device resident
        ms_g, st = timed(lambda: cd.decompress_ptr(d_comp.data_ptr(), 1, offs, rows, NC, 2, fl, d_out.data_ptr(), 1))
        tm_g = cd.timings()
        ln += cd.launches()
        assert not st.any()
        # (e) decompress reference-written streams end to end (host .cbin bytes -> host array)
        ms_re, st = timed(lambda: cd.decompress_ptr(h_ref_t.data_ptr(), 0, ref_offs_e, rows_e, NC, 2, fl, h_out_t.data_ptr(), 0))
        ln += cd.launches()
        assert not st.any()
        if record is not None:
            record.append(dict(c=ms, ce=ms_e, r=ms_r, g=ms_g, re=ms_re, tm=tm, tm_r=tm_r, tm_g=tm_g, launches=ln,
                               csize=int(offs[-1]), csize_e=int(offs_e[-1])))

    for _ in range(args.warmup):
        step(None)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    rec = []
    for _ in range(args.steps):
        step(rec)
    barrier()
    clocks = sampler.stop() if sampler else None

    # ---- verification outside the timed region: e2e decode == input, device decode == input
    assert np.array_equal(h_out_t.numpy(), h_raw), 'e2e decode of reference-written streams differs from the input'
    assert torch.equal(d_out, d_raw), 'device decode differs from the input'

    keys = ['c', 'ce', 'r', 'g', 're']
    mean_ms = np.array([np.mean([r[k] for r in rec]) for k in keys])
    mx = max_over_ranks(mean_ms)          # max over ranks of the per-step means
    gbps = {k: world * (e2e_bytes if k in ('ce', 're') else raw_bytes) / (mx[i] / 1e3) / 1e9 for i, k in enumerate(keys)}
    # ---- copy-only ceiling of the host link with all ranks copying at once, and the extra legs
    sys.path.insert(0, str(ROOT / 'tools'))
    import bench_legs
    del h_comp_t, h_out_t, h_ref_t
    ceil_bytes = min(e2e_bytes, 4 << 30)
    h2d_s, d2h_s = max_over_ranks(bench_legs.copy_ceiling(ceil_bytes, barrier))
    legs = {}
    if not args.no_legs:
        del d_raw, d_comp, d_out, d_ref, h_raw_t
        torch.cuda.empty_cache()
        legs['sharded_file'] = bench_legs.sharded_leg(rank, world, args.shard_chunks)
        if world == 1:
            legs['lfp_1s_chunks'] = bench_legs.lfp_leg(cd, args.lfp_chunks, 2500, threads)
            legs['lfp_0.1s_chunks'] = bench_legs.lfp_leg(cd, args.lfp_small_chunks, 250, threads)
            legs['latency'] = bench_legs.latency_leg(threads, n_chunks=args.latency_chunks)
            legs['file_to_file'] = bench_legs.file_leg()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    csize = rec[-1]['csize']
    ref_total = int(ref_offs[-1])
    hbm, how = peaks()
    lz_ms = float(np.mean([r['tm'][3] for r in rec]))
    alg_bytes = raw_bytes + csize               # read raw + write compressed = (1 + r) B per raw byte
    achieved = alg_bytes / (lz_ms / 1e3) / 1e9
    prof = lz77_profile()
    traffic = prof.get('dram_bytes_per_raw_byte')
    traffic = traffic * raw_bytes if traffic else None
    inf_ms = float(np.mean([r['tm_r'][2] for r in rec]))
    # the HBM-bound stages of the path against the same measured peak (algorithmic bytes: transform reads and writes
    # every byte once; adler32 reads it once; the single-pass inverse reads T once and writes the output once)
    def hbm_stage(ms, bytes_per_raw):
        a = bytes_per_raw * raw_bytes / (ms / 1e3) / 1e9
        return {'ms': ms, 'achieved': a, 'frac': a / hbm, 'bytes_per_raw_byte': bytes_per_raw}
    tr_ms = float(np.mean([r['tm'][1] for r in rec]))
    inv_ms = float(np.mean([r['tm_r'][4] for r in rec]))
    hbm_stages = {'fwd_tile_kernel': hbm_stage(tr_ms, 2), 'inv_tile_kernel': hbm_stage(inv_ms, 2)}
    ad_ms = float(np.mean([r['tm_r'][3] for r in rec]))
    if ad_ms > 0:
        hbm_stages['adler_partial_kernel (decode)'] = hbm_stage(ad_ms, 1)
    dec_names = ['h2d', 'plan', 'inflate', 'adler', 'inverse', 'd2h', '-', 'total']
    stage_r = {k: float(np.mean([r['tm_r'][i] for r in rec])) for i, k in enumerate(dec_names) if k != '-'}
    stage_g = {k: float(np.mean([r['tm_g'][i] for r in rec])) for i, k in enumerate(dec_names) if k != '-'}
    line = {
        'metric': METRIC, 'value': gbps['c'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': float(mx[0]), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'int16', 'data': 'synthetic',
        'config': workload_config(n_chunks, n_distinct),
        'run': {'seg_bytes': cd.get_param('seg_bytes'), 'e2e_chunks_per_gpu': n_e2e,
                'lz_ctas_per_sm': cd.get_param('lz_ctas_per_sm')},
        'e2e': {'value': gbps['ce'], 'unit': UNIT, 'h2d_bytes_per_step': e2e_bytes, 'd2h_bytes_per_step': rec[-1]['csize_e']},
        'e2e_ceiling': {
            'h2d_GBps_all_ranks': world * ceil_bytes / h2d_s / 1e9, 'd2h_GBps_all_ranks': world * ceil_bytes / d2h_s / 1e9,
            'compress_GBps': world * ceil_bytes / max(h2d_s, d2h_s * rec[-1]['csize_e'] / e2e_bytes) / 1e9,
            'decompress_GBps': world * ceil_bytes / max(d2h_s, h2d_s * rec[-1]['csize_e'] / e2e_bytes) / 1e9,
            'note': 'copy-only pinned-memory transfers of %.1f GB per rank, all ranks at once (max over ranks): what the '
                    'host link alone allows for raw-in/compressed-out and compressed-in/raw-out' % (ceil_bytes / 1e9)},
        'configs': legs,
        'decompress': {
            'roofline': {'kernel': 'block-parallel inflate (par_find/validate/block/lz), reference-written streams',
                         'bound': 'issue', 'achieved': (raw_bytes + ref_total) / (inf_ms / 1e3) / 1e9, 'peak': hbm,
                         'unit': 'GB/s', 'frac': (raw_bytes + ref_total) / (inf_ms / 1e3) / 1e9 / hbm,
                         'note': 'algorithmic bytes = compressed read + raw written per step, over the inflate stage'},
            'reference_written': {'value': gbps['r'], 'unit': UNIT, 'e2e_value': gbps['re'], 'inflate_ms': inf_ms,
                                  'streams': n_chunks, 'stage_ms': stage_r,
                                  'note': 'index-less zlib streams (what the reference Writer emits): block-parallel decoder'},
            'gpu_written': {'value': gbps['g'], 'unit': UNIT, 'stage_ms': stage_g, 'note': 'in-band index of segments and sub-blocks + the encoder\'s step rule: seg_tokens_kernel (lane per sub-block) and seg_resolve_kernel (all tokens of a step at once)'}},
        'ratio': {'gpu_comp_over_raw': csize / raw_bytes, 'zlib6_comp_over_raw': ref_total / raw_bytes,
                  'gpu_size_over_zlib': csize / ref_total, 'north_star_limit': 1.031},
        'roofline': {'kernel': 'lz77_kernel<2,512>', 'bound': 'issue', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s',
                     'frac': achieved / hbm, 'traffic': traffic, 'peak_source': how,
                     'issue_slot_frac': prof.get('issue_slot_frac'), 'alu_pipe_frac': prof.get('alu_pipe_frac'),
                     'warp_inst_per_raw_byte': prof.get('warp_inst_per_raw_byte'),
                     'note': 'frac = algorithmic bytes (raw + compressed per step) / kernel time / measured HBM peak, as '
                             'the contract defines it; the kernel is bound by instruction issue (integer ALU pipe), '
                             'not by HBM: issue_slot_frac / alu_pipe_frac are from the committed ncu capture '
                             '(profiles/r02_lz77_ncu.json)',
                     'hbm_bound_stages': hbm_stages,
                     'stage_ms': dict(zip(['h2d', 'transform', 'adler', 'lz77', 'huffman_scan', 'encode', 'd2h', 'total'],
                                          [float(np.mean([r['tm'][i] for r in rec])) for i in range(8)]))},
        'cpu_baseline': {'value': cpu_c, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'decompress_value': cpu_d,
                         'sample': '%d chunks (%d distinct) through NumPy diff + zlib level 6 in a %d-thread pool, '
                                   '%.1f s' % (n_cpu, n_distinct, threads, cpu_secs)},
        'gpu_launches': int(sum(r['launches'] for r in rec)),
        'clocks': clocks,
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
