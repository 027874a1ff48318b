"""Pins for the ebur128 / loudnorm restatement (no ffmpeg binary here: parity vs ffmpeg is unpinned;
these are the external known answers of SURVEY.md section 4)."""
import math

import numpy as np

from oracle import chain, cport


def test_bs1770_48k_coefficient_table():
    """ITU-R BS.1770-4 Table 1/2 (48 kHz)."""
    _, _, (pb, pa, rb, ra) = chain.k_weighting(48000)
    assert np.allclose(pb, [1.53512485958697, -2.69169618940638, 1.19839281085285], atol=1e-13)
    assert np.allclose(pa, [1.0, -1.69065929318241, 0.73248077421585], atol=1e-13)
    assert np.allclose(rb, [1.0, -2.0, 1.0])
    assert np.allclose(ra, [1.0, -1.99004745483398, 0.99007225036621], atol=1e-13)


def _sine(fs, secs, freq, dbfs, stereo=True):
    t = np.arange(int(fs * secs)) / fs
    s = np.round(32767 * 10 ** (dbfs / 20) * np.sin(2 * np.pi * freq * t)).astype(np.int16)
    z = np.zeros_like(s)
    return np.stack([s, s if stereo else z], axis=1)


def test_ebu_3341_case1_stereo_1k_minus23():
    assert abs(chain.integrated_loudness(_sine(48000, 20, 1000, -23.0), 48000) - (-23.0)) < 0.1


def test_997hz_full_scale_mono_is_minus_3_01():
    # ffmpeg's ebur128 is histogram-only: a steady tone reports its 0.1 LU bin centre (-3.05 for -3.01)
    v = chain.integrated_loudness(_sine(48000, 10, 997, 0.0, stereo=False), 48000)
    assert abs(v - (-3.01)) <= 0.05 and abs(v - (-3.05)) < 1e-9


def test_relative_gate_ignores_quiet_passage():
    loud = _sine(48000, 10, 1000, -20.0)
    quiet = _sine(48000, 10, 1000, -45.0)   # > -70 LUFS absolute gate, < relative gate
    both = np.concatenate([loud, quiet, loud])
    # transition blocks pass the gate and pull ~0.07 LU; an ungated mean would sit near -21.7
    assert abs(chain.integrated_loudness(both, 48000) - chain.integrated_loudness(loud, 48000)) < 0.1


def test_silence_is_minus_inf_and_copied_through():
    z = np.zeros((48000, 2), dtype=np.int16)
    info = {}
    out = chain.normalize(z, 48000, -14.0, info)
    assert info["input_i"] == -math.inf and not info["normalized"]
    assert np.array_equal(out, z)
    short = _sine(48000, 0.3, 1000, -10.0)  # < 400 ms: no gating block at all
    assert chain.integrated_loudness(short, 48000) == -math.inf


def test_literal_df2_matches_lfilter_restatement():
    b, a, _ = chain.k_weighting(44100)
    x = _sine(44100, 1.0, 440, -6.0)
    lit = cport.kfilter_df2(x, b, a)
    assert np.max(np.abs(lit - chain.k_weighted(x, 44100, literal=False))) < 1e-9


def test_independent_crosscheck_torchaudio():
    """torchaudio's BS.1770 loudness (non-histogram) is an independent implementation."""
    import torch
    import torchaudio
    from audio_mastering_engine_b200 import synth
    x = synth.track(6.0, 48000, track_id=2)
    ours = chain.integrated_loudness(x, 48000)
    wav = torch.from_numpy(x.astype(np.float32).T / 32768.0)
    theirs = float(torchaudio.functional.loudness(wav, 48000))
    assert abs(ours - theirs) < 0.1


def test_static_gain_rounding_and_apply():
    g, mi = chain.static_gain_from_measured(-20.126, -14.0)
    assert mi == -20.13 and g == math.pow(10.0, (-14.0 + 20.13) / 20.0)
    x = np.array([[100, -100], [32767, -32768], [3, -3]], dtype=np.int16)
    y = chain.apply_static_gain(x, 1.5)
    assert y.tolist() == [[150, -150], [32767, -32768], [4, -4]]  # lrint: 4.5 -> 4 (half to even)


def test_loudness_range_known_answer():
    """EBU Tech 3342 style: equal halves at -20 and -30 LUFS => LRA = 10 LU (+-1)."""
    a = _sine(48000, 20, 1000, -20.0)
    b = _sine(48000, 20, 1000, -30.0)
    x = np.concatenate([a, b])
    _, sub = chain.gating_block_energies(x, 48000)
    lra = chain.loudness_range_from_histogram(chain.short_term_histogram(sub, 48000))
    assert abs(lra - 10.0) <= 1.0
    _, sub = chain.gating_block_energies(a, 48000)
    assert chain.loudness_range_from_histogram(chain.short_term_histogram(sub, 48000)) < 0.2
    assert chain.loudness_range_from_histogram(np.zeros(1000, dtype=np.int64)) == 0.0


def test_block_energy_summation_order_does_not_move_a_histogram_bin():
    """libebur128 (ebur128_calc_gating_block) adds the 400 ms of squares in ONE loop per channel; the restatement and the
    kernels add four 100 ms sums.  Same value to FP64 round-off, and the 0.1 LU histogram bins are ~2.3 % wide: on the
    synthetic material the two orders give the SAME histogram (hence the same integrated loudness, bit for bit)."""
    from audio_mastering_engine_b200 import synth
    fs = 48000
    s100 = chain.samples_in_100ms(fs)
    for tid in (0, 1, 5):
        x = synth.track(8.0, fs, track_id=tid)
        blocks, _ = chain.gating_block_energies(x, fs)
        y = chain.k_weighted(x, fs)
        direct = np.empty(len(blocks))
        for i in range(len(blocks)):
            seg = y[i * s100:(i + 4) * s100]
            tot = 0.0
            for c in range(seg.shape[1]):                 # per channel, sequentially, as the C loop does
                acc = 0.0
                for v in seg[:, c].tolist():
                    acc += v * v
                tot += acc
            direct[i] = tot / float(4 * s100)
        assert np.max(np.abs(direct - blocks) / blocks) < 1e-12
        assert np.array_equal(chain.block_histogram(direct), chain.block_histogram(blocks))
