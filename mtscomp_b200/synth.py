# -*- coding: utf-8 -*-
"""Seeded synthetic Neuropixels-like recordings (SURVEY.md §8d generators).

These are the workloads of BASELINE.json's configs: an AP-band generator (band-limited noise, rms 9 counts,
injected biphasic spikes, a sync square wave on the last channel) and an LFP-band generator (a few smooth latent
sources mixed across channels).  Host-side, NumPy/SciPy only; used by bench.py, the tests and the golden-vector script.
"""

import numpy as np


def _sync_channel(ns, sample_rate, t0=0):
    # 0/64 square wave toggling every 0.5 s
    t = (np.arange(ns) + t0) / float(sample_rate)
    return (np.floor(t / 0.5).astype(np.int64) % 2 * 64).astype(np.int16)


def ap_chunk(ns=30000, nc=385, sample_rate=30000., seed=1234, rms=9.0, t0=0):
    """One AP-band chunk, shape (ns, nc) int16, values within [-512, 511]."""
    from scipy.signal import butter, sosfilt
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((ns, nc), dtype=np.float32)
    hi = min(6000., 0.45 * sample_rate)
    lo = min(300., 0.05 * sample_rate)
    sos = butter(3, [lo, hi], btype='band', fs=sample_rate, output='sos')
    x = sosfilt(sos, x, axis=0).astype(np.float32)
    x *= rms / max(float(np.sqrt(np.mean(x * x))), 1e-9)
    # spikes: 8 Hz per 8-channel group, biphasic 1.5 ms template, amplitude U(40,200), sigma 1.5 ch over +-4 ch
    tl = max(int(round(1.5e-3 * sample_rate)), 4)
    tt = np.arange(tl, dtype=np.float32) * (45.0 / tl)
    templ = -np.exp(-((tt - 12) / 4) ** 2) + 0.4 * np.exp(-((tt - 24) / 8) ** 2)
    n_groups = max(nc // 8, 1)
    n_spk = rng.poisson(8.0 * ns / sample_rate * n_groups)
    if ns > tl + 1:
        times = rng.integers(0, ns - tl, n_spk)
        chans = rng.integers(0, nc, n_spk)
        amps = rng.uniform(40, 200, n_spk)
        for t, c, a in zip(times, chans, amps):
            c0, c1 = max(c - 4, 0), min(c + 5, nc)
            prof = np.exp(-0.5 * ((np.arange(c0, c1) - c) / 1.5) ** 2).astype(np.float32)
            x[t:t + tl, c0:c1] += a * templ[:, None] * prof[None, :]
    out = np.ascontiguousarray(np.clip(np.rint(x), -512, 511).astype(np.int16))
    out[:, nc - 1] = _sync_channel(ns, sample_rate, t0)
    return out


def lfp_chunk(ns=2500, nc=385, sample_rate=2500., seed=4321, t0=0):
    """One LFP-band chunk: 12 latent band-limited sources with Gaussian channel profiles + white noise."""
    from scipy.signal import butter, sosfilt
    rng = np.random.default_rng(seed)
    n_src = 12
    src = rng.standard_normal((ns + 2000, n_src)).astype(np.float32)
    hi = min(300., 0.45 * sample_rate)
    sos = butter(2, [0.5, hi], btype='band', fs=sample_rate, output='sos')
    src = sosfilt(sos, src, axis=0)[2000:].astype(np.float32)
    centers = rng.uniform(0, nc, n_src)
    prof = np.exp(-0.5 * ((np.arange(nc)[None, :] - centers[:, None]) / 40.0) ** 2).astype(np.float32)
    x = src @ prof
    x *= 60.0 / max(float(np.sqrt(np.mean(x * x))), 1e-9)
    x += 6.0 * rng.standard_normal((ns, nc), dtype=np.float32)
    out = np.ascontiguousarray(np.clip(np.rint(x), -512, 511).astype(np.int16))
    out[:, nc - 1] = _sync_channel(ns, sample_rate, t0)
    return out


def ap_recording(n_chunks, ns=30000, nc=385, sample_rate=30000., seed=1234, n_distinct=None):
    """(n_chunks*ns, nc) int16 built from `n_distinct` seeded base chunks tiled in order (chunks are independent
    in the codec, so tiling only bounds the generation time; the bench states n_distinct in its config)."""
    n_distinct = min(n_distinct or n_chunks, n_chunks)
    base = [ap_chunk(ns, nc, sample_rate, seed + i, t0=i * ns) for i in range(n_distinct)]
    return np.concatenate([base[i % n_distinct] for i in range(n_chunks)], axis=0)
