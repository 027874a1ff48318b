"""Synthetic int16 stereo inputs for the BASELINE.json configs.

numpy generators are the specification (used by the parity tests and the golden fixtures);
``torch_track_batch`` is the same recipe evaluated with torch so ``bench.py`` can fill a
multi-gigabyte batch in seconds (it is a different random stream, not a bit-copy of the numpy one).
"""
from __future__ import annotations

import numpy as np

DEFAULT_MULTIBAND = {"low_thresh": -25.0, "low_ratio": 6.0, "mid_thresh": -20.0, "mid_ratio": 3.0,
                     "high_thresh": -15.0, "high_ratio": 4.0}  # GUI defaults, mastering_gui.py:50-52

ALL_BOOST_EQ = {"bass_boost": 2.0, "mid_cut": 1.0, "presence_boost": 1.5, "treble_boost": 1.0}


def _pink(rng, n):
    w = rng.standard_normal(n)
    spec = np.fft.rfft(w)
    f = np.arange(spec.shape[0], dtype=np.float64)
    f[0] = 1.0
    spec = spec / np.sqrt(f)
    spec[0] = 0.0
    p = np.fft.irfft(spec, n)
    rms = np.sqrt(np.mean(p * p))
    return p * (10 ** (-26.0 / 20.0) / rms)


def base_signal(n, fs, track_id=0):
    """float64 [n,2] in roughly [-0.7, 0.7]: correlated pink noise + 5 sines, before quantising."""
    rng = np.random.default_rng(20260 + int(track_id))
    pa, pb = _pink(rng, n), _pink(rng, n)
    left, right = pa + 0.5 * pb, pa - 0.5 * pb
    t = np.arange(n, dtype=np.float64) / fs
    amp = 10 ** (-30.0 / 20.0)
    for f in (55.0, 220.0, 1000.0, 4000.0, 9000.0):
        left = left + amp * np.sin(2 * np.pi * f * t)
        right = right + amp * np.cos(2 * np.pi * f * t)
    return np.stack([left, right], axis=1)


def quantise(x):
    x = 0.7 * np.tanh(x / 0.7)
    return np.round(x * 32767.0).astype(np.int16)


def track(seconds, fs, track_id=0, am_hz=None, am_db=6.0, drift_db=None, drift_period=90.0):
    n = int(round(seconds * fs))
    x = base_signal(n, fs, track_id)
    t = np.arange(n, dtype=np.float64) / fs
    if am_hz:
        x = x * (10 ** ((am_db * np.sin(2 * np.pi * am_hz * t)) / 20.0))[:, None]
    if drift_db:
        x = x * (10 ** ((drift_db * np.sin(2 * np.pi * t / drift_period)) / 20.0))[:, None]
    return quantise(x)


def stress_track(seconds, fs, track_id=0):
    """C5: 10 s sections alternating -12 dBFS bursts / -45 dBFS beds, exact-zero gaps of 3 s at the
    start of every bed, single-sample +-0.9 clicks every 0.5 s."""
    n = int(round(seconds * fs))
    x = base_signal(n, fs, track_id)
    rms = np.sqrt(np.mean(x * x))
    t = np.arange(n, dtype=np.float64) / fs
    sec = (t // 10.0).astype(np.int64)
    loud = (sec % 2 == 0)
    level = np.where(loud, 10 ** (-12.0 / 20.0), 10 ** (-45.0 / 20.0)) / rms
    x = x * level[:, None]
    gap = (~loud) & ((t % 10.0) < 3.0)
    x[gap] = 0.0
    pcm = quantise(x)
    clicks = np.arange(0, n, int(0.5 * fs))
    sign = np.where((np.arange(len(clicks)) % 2) == 0, 1, -1)
    v = (sign * int(round(0.9 * 32767))).astype(np.int16)
    pcm[clicks, 0] = v
    pcm[clicks, 1] = v
    return pcm


def c1_settings():
    return dict(ALL_BOOST_EQ, analog_character=25, width=1.2, lufs=-14.0, multiband=False)


def c2_settings():
    return dict(ALL_BOOST_EQ, analog_character=25, width=1.2, lufs=-14.0, multiband=True, **DEFAULT_MULTIBAND)


def c4_settings(track_id, presets):
    """Deterministic settings sweep of config C4; ``presets`` = EQ_PRESETS."""
    t = int(track_id)
    names = [None] + list(presets.keys())
    name = names[t % 6]
    s = {"bass_boost": 0.0, "mid_cut": 0.0, "presence_boost": 0.0, "treble_boost": 0.0}
    if name is not None:
        s.update(presets[name])
    s["width"] = [0.8, 1.0, 1.2][(t // 6) % 3]
    s["analog_character"] = [0, 25][(t // 18) % 2]
    s["lufs"] = [-16.0, -14.0, -9.0][(t // 36) % 3]
    s["multiband"] = (t % 2 == 1)
    s.update(DEFAULT_MULTIBAND)
    return s


def torch_track_batch(n_tracks, seconds, fs, device, first_track_id=0, am_hz=2.0, am_db=6.0):
    """[n_tracks, n, 2] int16 on ``device``: the base_signal recipe evaluated with torch."""
    import torch
    n = int(round(seconds * fs))
    out = torch.empty((n_tracks, n, 2), dtype=torch.int16, device=device)
    t = torch.arange(n, dtype=torch.float32, device=device) / fs
    f = torch.arange(n // 2 + 1, dtype=torch.float32, device=device)
    f[0] = 1.0
    inv = torch.rsqrt(f)
    inv[0] = 0.0
    amp = 10 ** (-30.0 / 20.0)
    sines_l = sum(amp * torch.sin(2 * np.pi * fr * t) for fr in (55.0, 220.0, 1000.0, 4000.0, 9000.0))
    sines_r = sum(amp * torch.cos(2 * np.pi * fr * t) for fr in (55.0, 220.0, 1000.0, 4000.0, 9000.0))
    mod = 10 ** ((am_db * torch.sin(2 * np.pi * am_hz * t)) / 20.0) if am_hz else None
    for k in range(n_tracks):
        g = torch.Generator(device=device)
        g.manual_seed(20260 + first_track_id + k)
        ps = []
        for _ in range(2):
            w = torch.randn(n, generator=g, device=device, dtype=torch.float32)
            p = torch.fft.irfft(torch.fft.rfft(w) * inv, n)
            p = p * (10 ** (-26.0 / 20.0) / p.square().mean().sqrt())
            ps.append(p)
        left = ps[0] + 0.5 * ps[1] + sines_l
        right = ps[0] - 0.5 * ps[1] + sines_r
        x = torch.stack([left, right], dim=1)
        if mod is not None:
            x = x * mod[:, None]
        x = 0.7 * torch.tanh(x / 0.7)
        out[k] = torch.round(x * 32767.0).to(torch.int16)
    return out
