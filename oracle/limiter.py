"""CPU restatement of ffmpeg's ``alimiter`` as the reference calls it (TEST INFRASTRUCTURE - see oracle/README.md).

Call site: ``audio_mastering_engine.py:223`` -
``ffmpeg -i normalized.wav -af alimiter=level_in=1:level_out=1:limit=0.98:attack=5:release=50 output``.

The ffmpeg binary and its sources are absent here and the reference does not pin a version (README.md:54-57):
this is restated from FFmpeg's ``libavfilter/af_alimiter.c`` (the Calf-derived look-ahead limiter, FFmpeg 5.x - 7.x)
[memory].  PARITY UNPINNED against ffmpeg itself; pinned only by the known answers in tests/test_oracle_limiter.py
(a signal that never exceeds the limit comes out delayed and scaled by exactly 1/limit; a single over-limit click is
met by an attack ramp that reaches limit/peak when the click leaves the look-ahead buffer, followed by a linear
release over `release` ms; nothing ever exceeds full scale).

What the filter does, per frame n (all channels linked), in double precision:
  * the frame enters a ring buffer of B = attack * fs frames; the frame that leaves it is the output frame, so the
    output is the input delayed by B - 1 frames (default ``latency=0``: the delay is not compensated, the stream
    keeps its length, the first B - 1 output frames are the zeros the buffer started with);
  * if the entering frame's peak exceeds ``limit`` a ramp is scheduled so that the attenuation is limit / peak at
    the moment that frame leaves the buffer: either the slope in force is replaced (a steeper one is needed now) or
    the peak is queued behind the peaks already scheduled (``nextpos`` / ``nextdelta``);
  * ``att += delta`` every frame; when a queued peak leaves the buffer the attenuation is set to limit / peak
    exactly and the slope becomes that peak's release slope (1 - limit / peak) / (fs * release) or the slope towards
    the next queued peak; back at 1.0 the state is reset;
  * ``asc`` (automatic release control) is OFF by default and the reference does not switch it on;
  * output = clip(x * att, -limit, limit) * (1 / limit) * level_out   (``level`` = auto level, ON by default: the
    file gets 1 / 0.98 louder even when nothing is limited);
  * s16 <-> double conversions around the filter are swresample's: x / 32768 in, clip(lrint(y * 32768)) out.
"""
from __future__ import annotations

import math

import numpy as np


def limiter_constants(fs, attack_ms=5.0, release_ms=50.0, channels=2):
    """buffer_size (samples, a multiple of `channels`), look-ahead frames B, release in seconds."""
    attack, release = attack_ms / 1000.0, release_ms / 1000.0
    buffer_size = int(fs * attack * channels)
    buffer_size -= buffer_size % channels
    return buffer_size, buffer_size // channels, release


def alimiter_py(pcm, fs, limit=0.98, attack_ms=5.0, release_ms=50.0, level_in=1.0, level_out=1.0, auto_level=True,
                return_att=False):
    """Literal per-frame restatement (asc off).  int16[N,2] -> int16[N,2]."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n_frames, channels = pcm.shape
    buffer_size, _, release = limiter_constants(fs, attack_ms, release_ms, channels)
    if buffer_size < channels:
        raise ValueError("attack too short for this sample rate")
    src = pcm.astype(np.float64).reshape(-1) * (1.0 / 32768.0)
    dst = np.zeros(n_frames * channels)
    buffer = [0.0] * buffer_size
    nextpos = [-1] * buffer_size
    nextdelta = [0.0] * buffer_size
    att, delta, pos, nextiter, nextlen = 1.0, 0.0, 0, 0, 0
    level = 1.0 / limit if auto_level else 1.0
    att_trace = np.zeros(n_frames) if return_att else None
    for n in range(n_frames):
        peak = 0.0
        for c in range(channels):
            sample = src[n * channels + c] * level_in
            buffer[pos + c] = sample
            peak = max(peak, abs(sample))
        if peak > limit:
            patt = min(limit / peak, 1.0)
            rdelta = (1.0 - patt) / (fs * release)
            d = (limit / peak - att) / buffer_size * channels
            found = False
            if d < delta:
                delta = d
                nextpos[0] = pos
                nextpos[1 % buffer_size] = -1
                nextdelta[0] = rdelta
                nextlen = 1
                nextiter = 0
            else:
                i = nextiter
                while i < nextiter + nextlen:
                    j = i % buffer_size
                    ppeak = 0.0
                    for c in range(channels):
                        ppeak = max(ppeak, abs(buffer[nextpos[j] + c]))
                    pdelta = (limit / peak - limit / ppeak) / (((buffer_size - nextpos[j] + pos) % buffer_size) / channels)
                    if pdelta < nextdelta[j]:
                        nextdelta[j] = pdelta
                        found = True
                        break
                    i += 1
                if found:
                    nextlen = i - nextiter + 1
                    nextpos[(nextiter + nextlen) % buffer_size] = pos
                    nextdelta[(nextiter + nextlen) % buffer_size] = rdelta
                    nextpos[(nextiter + nextlen + 1) % buffer_size] = -1
                    nextlen += 1
        b0 = (pos + channels) % buffer_size
        peak = 0.0
        for c in range(channels):
            peak = max(peak, abs(buffer[b0 + c]))
        att += delta
        for c in range(channels):
            dst[n * channels + c] = buffer[b0 + c] * att
        if b0 == nextpos[nextiter]:
            delta = nextdelta[nextiter]
            att = limit / peak
            nextlen -= 1
            nextpos[nextiter] = -1
            nextiter = (nextiter + 1) % buffer_size
        if att > 1.0:
            att, delta, nextiter, nextlen = 1.0, 0.0, 0, 0
            nextpos[0] = -1
        if att <= 0.0:
            att = 0.0000000000001
            delta = (1.0 - att) / (fs * release)
        if att != 1.0 and (1.0 - att) < 0.0000000000001:
            att = 1.0
        if delta != 0.0 and abs(delta) < 0.00000000000001:
            delta = 0.0
        for c in range(channels):
            v = dst[n * channels + c]
            v = min(max(v, -limit), limit)
            dst[n * channels + c] = v * level * level_out
        if return_att:
            att_trace[n] = att
        pos = (pos + channels) % buffer_size
    out = np.clip(np.rint(dst * 32768.0), -32768, 32767).astype(np.int16).reshape(n_frames, channels)
    return (out, att_trace) if return_att else out


def alimiter(pcm, fs, limit=0.98, attack_ms=5.0, release_ms=50.0):
    """The reference's call (level_in = level_out = 1, auto level on, asc off): C twin when built, else the Python loop."""
    from . import cport
    if cport.available():
        return cport.alimiter(pcm, fs, limit, attack_ms, release_ms)
    return alimiter_py(pcm, fs, limit, attack_ms, release_ms)
