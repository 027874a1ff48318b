"""Known answers for the alimiter restatement (oracle/limiter.py) and the true-peak meter (oracle/chain.py).
ffmpeg is absent and unpinned upstream (README.md:54-57): these pin the restatement to the behaviour af_alimiter.c
documents - look-ahead delay, auto level 1 / limit, attack ramp that meets limit / peak when the peak leaves the
look-ahead buffer, linear release - and the Python loop to its C twin, bit for bit."""
import numpy as np
import pytest

from audio_mastering_engine_b200 import synth
from oracle import chain, cport, limiter


def _scaled(x, g):
    return np.clip(np.rint(x.astype(np.float64) * g), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("fs", [22050, 44100, 48000, 96000])
@pytest.mark.parametrize("gain", [0.5, 4.5, 12.0])
def test_c_twin_equals_python_loop(fs, gain):
    assert cport.available()
    x = _scaled(synth.track(0.25, fs, track_id=5, am_hz=5.0), gain)
    a, att_a = limiter.alimiter_py(x, fs, return_att=True)
    b, att_b = cport.alimiter(x, fs, return_att=True)
    assert np.array_equal(a, b) and np.array_equal(att_a.view(np.int64), att_b.view(np.int64))
    assert np.abs(a.astype(np.int32)).max() <= 32768


def test_below_the_limit_is_delay_and_auto_level():
    """Nothing over 0.98: the output is the input delayed by B - 1 frames and scaled by exactly 1 / limit."""
    for fs in (44100, 48000, 192000):
        x = synth.track(0.2, fs, track_id=2) // 2
        assert np.abs(x).max() < 0.98 * 32768
        _, B, _ = limiter.limiter_constants(fs)
        assert B == {44100: 220, 48000: 240, 192000: 960}[fs]
        out = limiter.alimiter(x, fs)
        want = np.clip(np.rint(x.astype(np.float64) / 32768.0 * (1.0 / 0.98) * 32768.0), -32768, 32767).astype(np.int16)
        assert np.array_equal(out[B - 1:], want[: len(x) - (B - 1)])
        assert not out[: B - 1].any()


def test_single_click_attack_and_release():
    fs = 48000
    x = np.zeros((8000, 2), np.int16)
    x[:, 0] = 3000
    x[1000] = [32767, -20000]
    out, att = limiter.alimiter_py(x, fs, return_att=True)
    B, rel = 240, 2400
    peak = 32767 / 32768.0
    assert att[999] == 1.0                                     # nothing before the click enters the buffer
    ramp = att[1000:1000 + B - 1]
    assert np.all(np.diff(ramp) < 0) and np.allclose(np.diff(ramp), np.diff(ramp)[0], rtol=1e-9)   # linear attack
    assert att[1000 + B - 1] == 0.98 / peak                    # exactly limit / peak when the click leaves the buffer
    assert out[1000 + B - 1, 0] == 32767 and abs(int(out[1000 + B - 1, 1]) + 20000) <= 1   # 0.98 / 0.98 = full scale
    release = att[1000 + B - 1:1000 + B - 1 + rel]
    assert np.all(np.diff(release) > 0) and np.allclose(np.diff(release), (1 - 0.98 / peak) / rel, rtol=1e-9)
    assert np.all(att[1000 + B + rel:] == 1.0)                 # and back in the initial state


def test_true_peak_sees_between_the_samples():
    fs = 48000
    n = np.arange(4800)
    # fs / 4 sine sampled 45 degrees off its crests: sample peak = 0.5 / sqrt(2) ... true peak ~ 0.5
    x = 0.5 * np.sin(2 * np.pi * 0.25 * n + np.pi / 4)
    pcm = np.stack([np.rint(x * 32767), np.zeros_like(x)], axis=1).astype(np.int16)
    sample_peak = np.abs(pcm).max() / 32768.0
    tp = chain.true_peak(pcm, fs)
    assert sample_peak == pytest.approx(0.5 / np.sqrt(2), rel=1e-3)
    assert 0.47 < tp < 0.53
    assert chain.true_peak(pcm, 192000) == sample_peak         # no oversampling from 192 kHz on


def test_master_with_limiter_and_true_peak():
    fs = 44100
    x = synth.track(1.0, fs, track_id=9, am_hz=3.0)
    s = dict(synth.c1_settings(), lufs=-9.0)
    plain, _ = chain.master(x, fs, s)
    limited, info = chain.master(x, fs, dict(s, limiter=True, true_peak=True))
    assert np.array_equal(limited, limiter.alimiter(plain, fs))
    assert info["true_peak"] >= info["sample_peak"] / 32768.0 * 0.98
