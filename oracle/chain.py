"""CPU restatement of the reference mastering chain (TEST INFRASTRUCTURE - see oracle/README.md).

What is restated, and from where (all ``file:line`` are into the upstream reference tree,
``audio_mastering_engine.py`` unless noted):

* rows 2-9 of SURVEY.md section 8(a) - the int16<->float converters (:250-257), "analog character"
  warmth (:258-266), stereo width (:267-271), the 4-stage EQ (:272-298) and the 3-band split of the
  multiband compressor (:299-305).  These are numpy/scipy code in the reference; the restatement makes
  the SAME numpy/scipy calls on plain ``int16[N,2]`` arrays instead of pydub ``AudioSegment`` objects.
  Pinned: ``tests/golden/make_golden.py`` imports the reference's own module (pydub / ai_tagger
  stubbed) and the committed fixtures must match this file bit for bit (tests/test_oracle_golden.py).
* pydub ``compress_dynamic_range`` and ``AudioSegment.overlay`` (called at :306-309).  pydub is a
  third-party dependency that is NOT vendored in the reference and not installed here
  (requirements.txt:2, unpinned; latest release 0.25.1).  Restated from pydub 0.25.1 ``effects.py`` /
  ``audio_segment.py`` on top of CPython's real ``audioop`` (``rms``, ``mul``, ``add``).
  PARITY UNPINNED at this boundary: the reference has no tests or golden vectors for it.
* ffmpeg ``segment`` / ``concat`` (:178, :207-212) - in-memory slicing / concatenation.
* ffmpeg ``loudnorm`` two-pass in LINEAR mode + libavfilter ``ebur128.c`` integrated loudness
  (called at :227-246).  The ffmpeg binary is absent here and unpinned upstream (README.md:54-57).
  Restated from FFmpeg's ``libavfilter/ebur128.c`` and ``af_loudnorm.c``.  PARITY UNPINNED against
  ffmpeg itself; pinned against the ITU-R BS.1770-4 48 kHz coefficient table and EBU Tech 3341 style
  known answers (tests/test_oracle_loudness.py).

Declared deviations: D1 chunk length is a parameter (default 30*fs frames) instead
of ffmpeg's packet-granular cut; D2 pydub overlay's whole-millisecond slicing is not reproduced;
D3 loudnorm *dynamic* mode is not reproduced (static gain always); D4 loudness is measured at the
native rate, not ffmpeg's 192 kHz resampled stream, and the true peak is the BS.1770 Annex 2 meter;
the final ``alimiter`` (:223) is restated in oracle/limiter.py and applied when settings["limiter"] is set.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
from scipy.signal import butter, lfilter, sosfilt

with warnings.catch_warnings():
    warnings.simplefilter("ignore", DeprecationWarning)
    import audioop  # CPython <= 3.12 stdlib; the arithmetic pydub delegates to

# engine.py:32-38 - the preset table is part of the settings contract.
EQ_PRESETS = {
    "Vocal Clarity": {"bass_boost": -1.0, "mid_cut": 2.0, "presence_boost": 2.5, "treble_boost": 1.0},
    "Bass Punch": {"bass_boost": 2.5, "mid_cut": 1.0, "presence_boost": -1.0, "treble_boost": 0.5},
    "Vintage Warmth": {"bass_boost": 1.5, "mid_cut": 0.0, "presence_boost": -1.5, "treble_boost": -2.0},
    "Lo-Fi Haze": {"bass_boost": -2.0, "mid_cut": 3.0, "presence_boost": -2.0, "treble_boost": -4.0},
    "EDM Kick & Highs": {"bass_boost": 2.0, "mid_cut": 4.0, "presence_boost": 1.0, "treble_boost": 3.0},
}


# --------------------------------------------------------------------------------------------------
# converters - engine.py:250-257
# --------------------------------------------------------------------------------------------------
def to_float(pcm: np.ndarray) -> np.ndarray:
    """int16[N,2] -> float32[N,2] / 2**15   (engine.py:250-253)."""
    pcm = np.asarray(pcm, dtype=np.int16)
    return pcm.astype(np.float32) / (2 ** 15)


def to_pcm(x: np.ndarray) -> np.ndarray:
    """clip to [-1,1], times 32767 in the array's own dtype, truncate toward zero (engine.py:254-257)."""
    clipped = np.clip(x, -1.0, 1.0)
    return (clipped * 32767).astype(np.int16)


# --------------------------------------------------------------------------------------------------
# EQ stages - engine.py:272-298
# --------------------------------------------------------------------------------------------------
def shelf(samples, fs, cutoff_hz, gain_db, kind):
    """engine.py:283-289.  NB the cut branch is algebraically the bare Butterworth output."""
    if gain_db == 0.0:
        return samples
    b, a = butter(2, cutoff_hz / (0.5 * fs), btype=kind)
    y = lfilter(b, a, samples)  # default axis=-1, exactly as the reference calls it
    g = 10.0 ** (gain_db / 20.0)
    if gain_db > 0:
        return samples + (y - samples) * (g - 1)
    return samples * g + (y - samples * g)


def peak(samples, fs, center_hz, gain_db, q=1.41):
    """engine.py:290-298."""
    if gain_db == 0:
        return samples
    nyq = 0.5 * fs
    c = center_hz / nyq
    bw = c / q
    lo, hi = c - (bw / 2), c + (bw / 2)
    if lo <= 0:
        lo = 1e-9
    if hi >= 1.0:
        hi = 0.999999
    sos = butter(4, [lo, hi], btype="bandpass", output="sos")
    band = sosfilt(sos, samples)
    g = 10 ** (gain_db / 20.0)
    return samples + (band * (g - 1))


def eq_channel(ch, fs, settings):
    """engine.py:277-282 - fixed order low shelf, 1 kHz peak (-mid_cut), 4 kHz peak, high shelf."""
    ch = shelf(ch, fs, 250, settings.get("bass_boost", 0.0), "low")
    ch = peak(ch, fs, 1000, -settings.get("mid_cut", 0.0))
    ch = peak(ch, fs, 4000, settings.get("presence_boost", 0.0))
    ch = shelf(ch, fs, 8000, settings.get("treble_boost", 0.0), "high")
    return ch


def eq(samples, fs, settings):
    """engine.py:272-276 - per channel, written back IN PLACE into the float32 array."""
    if samples.ndim == 2:
        for i in range(samples.shape[1]):
            samples[:, i] = eq_channel(samples[:, i], fs, settings)
    else:
        samples = eq_channel(samples, fs, settings)
    return samples


# --------------------------------------------------------------------------------------------------
# warmth and width - engine.py:258-271
# --------------------------------------------------------------------------------------------------
def warmth(pcm, fs, percent):
    """engine.py:258-266.  The two shelf calls get the 2-D array, so lfilter runs ACROSS the two
    channels of each frame (axis=-1): a per-frame elementwise op, no recursion in time."""
    if percent == 0:
        return pcm
    cf = percent / 100.0
    x = to_float(pcm)
    drive = 1.0 + (cf * 0.5)
    sat = np.tanh(x * drive)
    sat = shelf(sat, fs, 120, cf * 1.0, "low")
    out = shelf(sat, fs, 12000, cf * 1.5, "high")
    return to_pcm(out)


def warmth_closed_form(pcm, fs, percent):
    """Same arithmetic as :func:`warmth` written per frame (what the CUDA kernel does).
    y0 = b0*L ; y1 = b0*R + (b1*L - a1*y0) ; blend x + (y-x)*(g-1), twice."""
    if percent == 0:
        return pcm
    cf = percent / 100.0
    x = to_float(pcm)
    drive = 1.0 + (cf * 0.5)
    s = np.tanh(x * drive)  # float32
    cur = s
    for fc, gdb, kind in ((120, cf * 1.0, "low"), (12000, cf * 1.5, "high")):
        b, a = butter(2, fc / (0.5 * fs), btype=kind)
        g = 10.0 ** (gdb / 20.0)
        c64 = cur.astype(np.float64)
        y0 = b[0] * c64[:, 0]
        z0 = (b[1] * c64[:, 0]) - a[1] * y0
        y1 = z0 + b[0] * c64[:, 1]
        y = np.stack([y0, y1], axis=1)
        cur = cur + (y - cur) * (g - 1)
    return to_pcm(cur)


def width(samples, w):
    """engine.py:267-271 (float32 in, float32 out)."""
    if samples.ndim != 2 or samples.shape[1] != 2:
        return samples
    left, right = samples[:, 0], samples[:, 1]
    mid, side = (left + right) / 2, (left - right) / 2
    side *= w
    nl, nr = np.clip(mid + side, -1.0, 1.0), np.clip(mid - side, -1.0, 1.0)
    return np.stack([nl, nr], axis=1)


# --------------------------------------------------------------------------------------------------
# pydub compress_dynamic_range / overlay restated on audioop  (call sites engine.py:306-309)
# --------------------------------------------------------------------------------------------------
def compressor_constants(fs, threshold, ratio, attack=5.0, release=50.0):
    """pydub effects.py: thresh_rms = max_possible_amplitude * db_to_float(threshold);
    frame_count(ms) = ms * (frame_rate / 1000.0) (a float); look_frames = int(frame_count(attack))."""
    thresh_rms = 32768.0 * (10 ** (float(threshold) / 20))
    attack_frames = attack * (fs / 1000.0)
    release_frames = release * (fs / 1000.0)
    look_frames = int(attack_frames)
    return thresh_rms, look_frames, attack_frames, release_frames


def max_attenuation_for_rms(rms, thresh_rms, ratio):
    """pydub: (1 - 1/ratio) * max(ratio_to_db(rms / thresh_rms), 0); ratio_to_db = 20*math.log(r, 10)."""
    if rms == 0:
        over = 0.0
    else:
        r = float(rms / thresh_rms)
        over = max(20 * math.log(r, 10), 0) if r != 0 else 0.0
    return (1 - (1.0 / ratio)) * over


def compress_dynamic_range_py(pcm, fs, threshold, ratio, attack=5.0, release=50.0, return_att=False):
    """Literal restatement of pydub 0.25.1 ``compress_dynamic_range`` (stereo-linked, per-frame Python
    loop) on real ``audioop``.  O(N * look_frames): use only on short inputs; the C twin
    (oracle/c/ame_oracle.c, checked bit-exact against this in tests) covers long ones."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n = pcm.shape[0]
    data = pcm.tobytes()
    fw = 4  # frame width: 2 channels * 2 bytes
    thresh_rms, look, attack_frames, release_frames = compressor_constants(fs, threshold, ratio, attack, release)
    out = []
    att_trace = np.zeros(n) if return_att else None
    att = 0.0
    for i in range(n):
        lo = max(i - look, 0)
        rms_now = audioop.rms(data[lo * fw:i * fw], 2)
        max_att = max_attenuation_for_rms(rms_now, thresh_rms, ratio)
        inc = max_att / attack_frames
        dec = max_att / release_frames
        if rms_now > thresh_rms and att <= max_att:
            att += inc
            att = min(att, max_att)
        else:
            att -= dec
            att = max(att, 0)
        frame = data[i * fw:(i + 1) * fw]
        if att != 0.0:
            frame = audioop.mul(frame, 2, 10 ** (-float(att) / 20))
        out.append(frame)
        if return_att:
            att_trace[i] = att
    res = np.frombuffer(b"".join(out), dtype=np.int16).reshape(-1, 2).copy()
    return (res, att_trace) if return_att else res


def overlay(a, b):
    """pydub ``a.overlay(b)`` for equal-length segments = audioop.add (saturating int16 add)."""
    a = np.ascontiguousarray(a, dtype=np.int16)
    b = np.ascontiguousarray(b, dtype=np.int16)
    return np.frombuffer(audioop.add(a.tobytes(), b.tobytes(), 2), dtype=np.int16).reshape(a.shape).copy()


def band_split(pcm, fs, low_crossover=250, high_crossover=4000):
    """engine.py:300-305 - Butterworth-4 LP/HP, mid by subtraction, each band truncated to int16."""
    x = to_float(pcm)
    low_sos = butter(4, low_crossover, btype="lowpass", fs=fs, output="sos")
    high_sos = butter(4, high_crossover, btype="highpass", fs=fs, output="sos")
    low, high = sosfilt(low_sos, x, axis=0), sosfilt(high_sos, x, axis=0)
    mid = x - low - high
    return to_pcm(low), to_pcm(mid), to_pcm(high)


def multiband(pcm, fs, settings, compress=None, taps=None):
    """engine.py:299-309."""
    compress = compress or compress_dynamic_range
    lo, mi, hi = band_split(pcm, fs)
    if taps is not None:
        taps["band_low"], taps["band_mid"], taps["band_high"] = lo, mi, hi
    lo_c = compress(lo, fs, settings.get("low_thresh"), settings.get("low_ratio"))
    mi_c = compress(mi, fs, settings.get("mid_thresh"), settings.get("mid_ratio"))
    hi_c = compress(hi, fs, settings.get("high_thresh"), settings.get("high_ratio"))
    if taps is not None:
        taps["comp_low"], taps["comp_mid"], taps["comp_high"] = lo_c, mi_c, hi_c
    return overlay(overlay(lo_c, mi_c), hi_c)


def compress_dynamic_range(pcm, fs, threshold, ratio, attack=5.0, release=50.0):
    """Dispatch: C twin when built (fast), literal Python loop otherwise."""
    from . import cport
    if cport.available():
        return cport.compress(pcm, fs, threshold, ratio, attack, release)
    return compress_dynamic_range_py(pcm, fs, threshold, ratio, attack, release)


# --------------------------------------------------------------------------------------------------
# one chunk of the hot loop - engine.py:189-197
# --------------------------------------------------------------------------------------------------
def process_chunk(pcm, fs, settings, taps=None, compress=None):
    """Stage order of engine.py:189-197 on one (already stereo, already 16-bit) chunk."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    if settings.get("analog_character", 0) > 0:
        pcm = warmth(pcm, fs, settings.get("analog_character"))
    if taps is not None:
        taps["warmth"] = pcm
    x = to_float(pcm)
    y = eq(x, fs, settings)
    if taps is not None:
        taps["eq"] = y.copy()
    if settings.get("width", 1.0) != 1.0:
        y = width(y, settings.get("width"))
    out = to_pcm(y)
    if taps is not None:
        taps["pre_multiband"] = out
    if settings.get("multiband"):
        out = multiband(out, fs, settings, compress=compress, taps=taps)
    return out


# --------------------------------------------------------------------------------------------------
# EBU R128 / BS.1770 integrated loudness, restated from FFmpeg libavfilter/ebur128.c
# --------------------------------------------------------------------------------------------------
def k_weighting(fs):
    """ebur128_init_filter(): pre-shelf and RLB high-pass, bilinear via K = tan(pi f0 / fs), the two
    biquads convolved into one 5-tap b / a pair.  Returns (b[5], a[5])."""
    f0, G, Q = 1681.974450955533, 3.999843853973347, 0.7071752369554196
    K = math.tan(math.pi * f0 / float(fs))
    Vh = math.pow(10.0, G / 20.0)
    Vb = math.pow(Vh, 0.4996667741545416)
    pb = [0.0, 0.0, 0.0]
    pa = [1.0, 0.0, 0.0]
    rb = [1.0, -2.0, 1.0]
    ra = [1.0, 0.0, 0.0]
    a0 = 1.0 + K / Q + K * K
    pb[0] = (Vh + Vb * K / Q + K * K) / a0
    pb[1] = 2.0 * (K * K - Vh) / a0
    pb[2] = (Vh - Vb * K / Q + K * K) / a0
    pa[1] = 2.0 * (K * K - 1.0) / a0
    pa[2] = (1.0 - K / Q + K * K) / a0
    f0, Q = 38.13547087602444, 0.5003270373238773
    K = math.tan(math.pi * f0 / float(fs))
    ra[1] = 2.0 * (K * K - 1.0) / (1.0 + K / Q + K * K)
    ra[2] = (1.0 - K / Q + K * K) / (1.0 + K / Q + K * K)
    b = [pb[0] * rb[0],
         pb[0] * rb[1] + pb[1] * rb[0],
         pb[0] * rb[2] + pb[1] * rb[1] + pb[2] * rb[0],
         pb[1] * rb[2] + pb[2] * rb[1],
         pb[2] * rb[2]]
    a = [pa[0] * ra[0],
         pa[0] * ra[1] + pa[1] * ra[0],
         pa[0] * ra[2] + pa[1] * ra[1] + pa[2] * ra[0],
         pa[1] * ra[2] + pa[2] * ra[1],
         pa[2] * ra[2]]
    return np.array(b), np.array(a), (np.array(pb), np.array(pa), np.array(rb), np.array(ra))


def histogram_tables():
    """ebur128.c init: 1000 bins of 0.1 LU from -70 to +30 LUFS."""
    bounds = np.empty(1001)
    energies = np.empty(1000)
    bounds[0] = math.pow(10.0, (-70.0 + 0.691) / 10.0)
    for i in range(1000):
        energies[i] = math.pow(10.0, (float(i) / 10.0 - 69.95 + 0.691) / 10.0)
    for i in range(1, 1001):
        bounds[i] = math.pow(10.0, (float(i) / 10.0 - 70.0 + 0.691) / 10.0)
    return bounds, energies


_BOUNDS, _ENERGIES = histogram_tables()


def find_histogram_index(energy, bounds=_BOUNDS):
    lo, hi = 0, 1000
    while True:
        mid = (lo + hi) // 2
        if energy >= bounds[mid]:
            lo = mid
        else:
            hi = mid
        if hi - lo == 1:
            return lo


def samples_in_100ms(fs):
    return (int(fs) + 5) // 10


def k_weighted(pcm, fs, literal=True):
    """s16 -> double (x / 32768, ffmpeg's s16->dbl) -> 4th-order K filter per channel.
    literal=True runs ebur128.c's own direct-form-II loop (C twin).  The lfilter form (DF-II
    transposed) is the same transfer function; the 4th-order direct forms are ill-conditioned near
    z=1, so the two differ by ~1e-11 (44.1 kHz) .. 1e-8 (192 kHz) of full scale - far inside 0.01 LU."""
    b, a, _ = k_weighting(fs)
    if literal:
        from . import cport
        if cport.available():
            return cport.kfilter_df2(pcm, b, a)
    x = np.asarray(pcm, dtype=np.int16).astype(np.float64) * (1.0 / 32768.0)
    return lfilter(b, a, x, axis=0)


def gating_block_energies(pcm, fs):
    """400 ms blocks every 100 ms: energy = sum_ch sum x^2 / block_frames  (ebur128_calc_gating_block).
    Returns (block energies, 100 ms sub-block energy sums per channel-summed)."""
    s100 = samples_in_100ms(fs)
    y = k_weighted(pcm, fs)
    n_sub = y.shape[0] // s100
    sq = (y[: n_sub * s100] ** 2).sum(axis=1)
    sub = sq.reshape(n_sub, s100).sum(axis=1) if n_sub else np.zeros(0)
    if n_sub < 4:
        return np.zeros(0), sub
    blocks = (sub[0:n_sub - 3] + sub[1:n_sub - 2] + sub[2:n_sub - 1] + sub[3:n_sub]) / float(4 * s100)
    return blocks, sub


def block_histogram(block_energies):
    hist = np.zeros(1000, dtype=np.int64)
    for e in block_energies:
        if e >= _BOUNDS[0]:
            hist[find_histogram_index(e)] += 1
    return hist


def gated_loudness_from_histogram(hist):
    """ebur128_gated_loudness(): absolute gate is the histogram floor (-70 LUFS), relative gate -10 LU.
    Returns (integrated LUFS or -inf, relative threshold energy)."""
    hist = np.asarray(hist, dtype=np.int64)
    count = int(hist.sum())
    if count == 0:
        return -math.inf, 0.0
    rel = 0.0
    for j in range(1000):
        rel += float(hist[j]) * _ENERGIES[j]
    rel /= float(count)
    rel *= math.pow(10.0, -10.0 / 10.0)
    if rel < _BOUNDS[0]:
        start = 0
    else:
        start = find_histogram_index(rel)
        if rel > _ENERGIES[start]:
            start += 1
    gated, above = 0.0, 0
    for j in range(start, 1000):
        gated += float(hist[j]) * _ENERGIES[j]
        above += int(hist[j])
    if above == 0:
        return -math.inf, rel
    gated /= float(above)
    return 10.0 * math.log10(gated) - 0.691, rel


def short_term_histogram(sub, fs):
    """ebur128.c LRA mode: 3 s windows (30 sub-blocks), the first ending at 3 s, then one every second."""
    s100 = samples_in_100ms(fs)
    hist = np.zeros(1000, dtype=np.int64)
    k = 0
    while 10 * k + 29 < len(sub):
        e = 0.0
        for j in range(30):
            e += float(sub[10 * k + j])
        e /= float(30 * s100)
        if e >= _BOUNDS[0]:
            hist[find_histogram_index(e)] += 1
        k += 1
    return hist


def loudness_range_from_histogram(hist):
    """ff_ebur128_loudness_range: blocks above (mean power - 20 dB), 10th to 95th percentile, in LU."""
    hist = np.asarray(hist, dtype=np.int64)
    n = int(hist.sum())
    if n == 0:
        return 0.0
    power = 0.0
    for j in range(1000):
        power += float(hist[j]) * _ENERGIES[j]
    power /= float(n)
    integ = math.pow(10.0, -20.0 / 10.0) * power
    if integ < _BOUNDS[0]:
        idx = 0
    else:
        idx = find_histogram_index(integ)
        if integ > _ENERGIES[idx]:
            idx += 1
    m = int(hist[idx:].sum())
    if m == 0:
        return 0.0
    p_lo, p_hi = int((m - 1) * 0.1 + 0.5), int((m - 1) * 0.95 + 0.5)
    acc, j = 0, idx
    while acc <= p_lo:
        acc += int(hist[j]); j += 1
    l_en = _ENERGIES[j - 1]
    while acc <= p_hi:
        acc += int(hist[j]); j += 1
    h_en = _ENERGIES[j - 1]
    return (10.0 * math.log10(h_en) - 0.691) - (10.0 * math.log10(l_en) - 0.691)


# ITU-R BS.1770-4 Annex 2: 4-phase x 12-tap interpolating FIR of the true-peak meter (the coefficient table of the
# recommendation).  y[4n + p] = sum_k h_p[k] x[n - k].
TRUE_PEAK_FIR = np.array([
    [0.0017089843750, 0.0109863281250, -0.0196533203125, 0.0332031250000, -0.0594482421875, 0.1373291015625,
     0.9721679687500, -0.1022949218750, 0.0476074218750, -0.0266113281250, 0.0148925781250, -0.0083007812500],
    [-0.0291748046875, 0.0292968750000, -0.0517578125000, 0.0891113281250, -0.1665039062500, 0.4650878906250,
     0.7797851562500, -0.2003173828125, 0.1015625000000, -0.0582275390625, 0.0330810546875, -0.0189208984375],
    [-0.0189208984375, 0.0330810546875, -0.0582275390625, 0.1015625000000, -0.2003173828125, 0.7797851562500,
     0.4650878906250, -0.1665039062500, 0.0891113281250, -0.0517578125000, 0.0292968750000, -0.0291748046875],
    [-0.0083007812500, 0.0148925781250, -0.0266113281250, 0.0476074218750, -0.1022949218750, 0.9721679687500,
     0.1373291015625, -0.0594482421875, 0.0332031250000, -0.0196533203125, 0.0109863281250, 0.0017089843750]])


def true_peak(pcm, fs):
    """BS.1770-4 Annex 2 true peak (linear, 1.0 = full scale): 4x oversampling below 96 kHz, 2x (phases 0 and 2) below
    192 kHz, the sample peak from there on.  NOT what ffmpeg's loudnorm prints as input_tp - that is the sample peak
    of the stream swresample'd to 192 kHz (deviation D4); both estimate the same inter-sample peak."""
    x = np.asarray(pcm, dtype=np.int16).astype(np.float64) * (1.0 / 32768.0)
    if x.shape[0] == 0:
        return 0.0
    if fs >= 192000:
        return float(np.abs(x).max())
    phases = (0, 1, 2, 3) if fs < 96000 else (0, 2)
    best = 0.0
    for c in range(x.shape[1]):
        for p in phases:
            y = np.convolve(x[:, c], TRUE_PEAK_FIR[p])[: x.shape[0]]
            best = max(best, float(np.abs(y).max()))
    return best


def integrated_loudness(pcm, fs):
    blocks, _ = gating_block_energies(pcm, fs)
    return gated_loudness_from_histogram(block_histogram(blocks))[0]


def static_gain_from_measured(measured_i, target_lufs):
    """engine.py:237-240 passes ``input_i`` as the '%.2f' string ffmpeg printed; af_loudnorm linear mode
    then applies pow(10, (target - measured_I) / 20)."""
    mi = float("%.2f" % measured_i)
    return math.pow(10.0, (float(target_lufs) - mi) / 20.0), mi


def apply_static_gain(pcm, gain):
    """s16 -> dbl (x/32768) -> * gain -> dbl -> s16 = clip(lrint(x * 32768))  (swresample conversions)."""
    x = np.asarray(pcm, dtype=np.int16).astype(np.float64) * (1.0 / 32768.0)
    y = x * gain
    return np.clip(np.rint(y * 32768.0), -32768, 32767).astype(np.int16)


def normalize(pcm, fs, target_lufs, info=None):
    """engine.py:227-246 in linear mode (deviations D3, D4)."""
    blocks, sub = gating_block_energies(pcm, fs)
    hist = block_histogram(blocks)
    measured, rel = gated_loudness_from_histogram(hist)
    if info is not None:
        info.update(input_lra=loudness_range_from_histogram(short_term_histogram(sub, fs)),
                    input_thresh=(10.0 * math.log10(rel) - 0.691) if hist.sum() else -70.0)
        info.update(input_i=measured, hist=hist, n_blocks=int(hist.sum()), rel_threshold_energy=rel,
                    sample_peak=int(np.abs(pcm.astype(np.int32)).max()) if len(pcm) else 0)
    if measured == -math.inf:  # engine.py:238-239 - silent audio, copy through
        if info is not None:
            info.update(gain=1.0, measured_i_2dp=-math.inf, normalized=False)
        return np.array(pcm, dtype=np.int16, copy=True)
    gain, mi = static_gain_from_measured(measured, target_lufs)
    if info is not None:
        info.update(gain=gain, measured_i_2dp=mi, normalized=True)
    return apply_static_gain(pcm, gain)


# --------------------------------------------------------------------------------------------------
# the whole path - engine.py:171-226 without the limiter (D5)
# --------------------------------------------------------------------------------------------------
def chunk_bounds(n_frames, fs, chunk_seconds=30):
    cf = int(chunk_seconds * fs)
    return [(s, min(s + cf, n_frames)) for s in range(0, n_frames, cf)]


def master(pcm, fs, settings, chunk_seconds=30, taps=None, compress=None):
    """split -> per-chunk chain (fresh state each chunk) -> concat -> normalise iff lufs is not None.
    Returns (int16[N,2], info)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    outs = []
    for s, e in chunk_bounds(pcm.shape[0], fs, chunk_seconds):
        outs.append(process_chunk(pcm[s:e], fs, settings, compress=compress))
    pre = np.concatenate(outs, axis=0) if outs else pcm.copy()
    info = {}
    if taps is not None:
        taps["pre_norm"] = pre
    if settings.get("lufs") is not None:
        out = normalize(pre, fs, settings.get("lufs"), info)
    else:
        out = pre
    if settings.get("true_peak"):
        info["true_peak"] = true_peak(pre, fs)
    if settings.get("limiter"):                            # engine.py:223, after the (optional) normalisation
        from . import limiter
        out = limiter.alimiter(out, fs, settings.get("limiter_limit", 0.98), settings.get("limiter_attack", 5.0),
                               settings.get("limiter_release", 50.0))
    return out, info
