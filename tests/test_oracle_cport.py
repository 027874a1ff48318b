"""C twin (oracle/c/ame_oracle.c) == literal pydub/audioop restatement, bit for bit."""
import numpy as np
import pytest

from oracle import chain, cport
from audio_mastering_engine_b200 import synth


@pytest.mark.parametrize("fs,thr,ratio", [(44100, -25.0, 6.0), (48000, -20.0, 3.0), (48000, -40.0, 10.0),
                                          (96000, -15.0, 4.0), (48000, 0.0, 1.0), (22050, -30.0, 2.5)])
def test_compressor_c_equals_python(fs, thr, ratio):
    assert cport.available()
    x = synth.track(0.12, fs, track_id=3, am_hz=25.0, am_db=10.0)
    x = (x.astype(np.int32) * 2).clip(-32768, 32767).astype(np.int16)
    x[2000:3500] = 0
    x[10] = [32767, -32768]
    py, att_py = chain.compress_dynamic_range_py(x, fs, thr, ratio, return_att=True)
    c, att_c = cport.compress(x, fs, thr, ratio, return_att=True)
    assert np.array_equal(att_py, att_c)
    assert np.array_equal(py, c)


def test_never_releases_below_threshold():
    """pydub quirk: attenuation is frozen while the level is under threshold."""
    fs = 48000
    x = np.zeros((fs // 2, 2), dtype=np.int16)
    t = np.arange(2400)
    x[1000:3400, 0] = x[1000:3400, 1] = (20000 * np.sin(2 * np.pi * 1000 * t / fs)).astype(np.int16)
    x[3400:] = 50  # quiet but non-zero tail
    _, att = cport.compress(x, fs, -20.0, 4.0, return_att=True)
    assert att[3400 + 300] > 1.0
    assert att[-1] == att[3400 + 300]


def test_window_rms_matches_audioop():
    import audioop
    x = synth.track(0.05, 48000, track_id=1)
    look = 240
    r = cport.window_rms(x, look)
    data = x.tobytes()
    for i in (0, 1, 5, 239, 240, 241, 1000, len(x) - 1):
        lo = max(i - look, 0)
        assert r[i] == audioop.rms(data[lo * 4:i * 4], 2)
