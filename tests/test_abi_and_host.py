"""CPU-side checks: the C-ABI library loads and exports every symbol include/ame.h declares (no compute
calls - there is no GPU here), the ctypes mirror matches the C structs, and the host-side design code
reproduces the reference's coefficient recipe."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from audio_mastering_engine_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from audio_mastering_engine_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ame.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ame_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ame.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.ame_abi_version() == _lib.AME_ABI_VERSION == 5


def test_struct_layout_matches(lib):
    import ctypes
    from audio_mastering_engine_b200 import _lib
    assert lib.ame_sizeof_track_params() == ctypes.sizeof(_lib.TrackParams)
    assert lib.ame_sizeof_track_result() == ctypes.sizeof(_lib.TrackResult)
    assert lib.ame_sizeof_plan_options() == ctypes.sizeof(_lib.PlanOptions)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from audio_mastering_engine_b200 import master, synth, AmeError
    with pytest.raises(AmeError):
        master(synth.track(0.05, 48000), 48000, synth.c1_settings())


def test_design_matches_reference_recipe():
    from scipy.signal import butter
    from audio_mastering_engine_b200 import design, _lib
    s = {"bass_boost": 2.0, "mid_cut": 1.0, "presence_boost": -1.5, "treble_boost": -4.0, "width": 1.2,
         "analog_character": 25, "lufs": -14, "multiband": True, "low_thresh": -25, "low_ratio": 6,
         "mid_thresh": -20, "mid_ratio": 3, "high_thresh": -15, "high_ratio": 4}
    lut = {}
    p = design.track_params(s, 44100, 1000, 0, 30, lut)
    assert p.flags == 15 and p.chunk_frames == 1323000 and lut == {25.0: 0}
    b, a = butter(2, 250 / (0.5 * 44100), btype="low")
    assert (p.eq[0].kind, p.eq[0].s[0].b0, p.eq[0].s[0].a2) == (_lib.AME_EQ_SHELF_BOOST, b[0], a[2])
    assert p.eq[0].gm1 == 10.0 ** (2.0 / 20.0) - 1
    assert p.eq[1].kind == _lib.AME_EQ_PEAK and p.eq[1].gm1 == 10 ** (-1.0 / 20.0) - 1
    assert p.eq[3].kind == _lib.AME_EQ_SHELF_CUT
    sos = butter(4, 250, btype="lowpass", fs=44100, output="sos")
    assert p.xlp[1].a1 == sos[1][4]
    assert p.comp[0].look_frames == 220 and p.comp[0].attack_frames == 220.5
    assert p.comp[1].thresh_rms == 32768.0 * (10 ** (-20.0 / 20))
    assert p.warm_eq > 1000 and p.warm_xover > 1000 and p.warm_kw > 3000
    flat = design.track_params({"lufs": None}, 48000, 10, 0, 30, {})
    assert flat.flags == 0 and all(flat.eq[i].kind == 0 for i in range(4)) and flat.warm_eq == 0
    with pytest.raises(ValueError, match="unsupported sample rate"):
        design.track_params({"lufs": -14.0}, 1000, 10, 0, 30, {})
    with pytest.raises(TypeError):        # multiband without its thresholds: pydub would raise too
        design.track_params({"multiband": True}, 48000, 10, 0, 30, {})


def test_bind_host_to_gpu_numa_is_harmless_without_topology():
    """No GPU / no sysfs topology: returns None and leaves the CPU affinity alone."""
    import os
    from audio_mastering_engine_b200 import bind_host_to_gpu_numa
    before = os.sched_getaffinity(0)
    node = bind_host_to_gpu_numa(0)
    assert node is None or isinstance(node, int)
    if node is None:
        assert os.sched_getaffinity(0) == before


def test_k_weighting_biquads_are_bs1770():
    from audio_mastering_engine_b200 import design
    (pb, pa), (rb, ra) = design.k_weighting_biquads(48000)
    assert np.allclose(pb, [1.53512485958697, -2.69169618940638, 1.19839281085285], atol=1e-13)
    assert np.allclose(ra, [1.0, -1.99004745483398, 0.99007225036621], atol=1e-13)


def test_warm_lut_is_numpy_tanh():
    from audio_mastering_engine_b200 import design
    lut = design.warm_lut(25)
    x = np.array([-32768, -1, 0, 1, 12345, 32767], dtype=np.int16)
    want = np.tanh((x.astype(np.float32) / 2 ** 15) * (1.0 + 0.25 * 0.5))
    assert lut.dtype == np.float32 and np.array_equal(lut[x.astype(np.int64) + 32768], want)


def test_warmup_length_bounds_state_error():
    """The tile warm-up must leave < 1e-12 relative error (host-side check with scipy)."""
    from scipy.signal import sosfilt
    from audio_mastering_engine_b200 import design
    fs = 48000
    w = design.eq_warm_frames(fs, 2.0, 1.0, 1.5, 1.0)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(3 * w)
    sos = design._peak_sos(fs, 1000)
    full = sosfilt(sos, x)
    cold = sosfilt(sos, x[w:])           # zero state at frame w
    err = np.abs(full[2 * w:] - cold[w:]).max() / np.sqrt(np.mean(full ** 2))
    assert err < 1e-12


def test_c_example_of_the_by_time_path_compiles(tmp_path):
    """examples/time_shard_nccl.c shows a non-Python host running one rank of the time-sharded path on include/ame.h
    alone (ame_shard_halo_exchange + ame_hist_allreduce): it must keep compiling against the header."""
    import subprocess
    src = os.path.join(ROOT, "examples", "time_shard_nccl.c")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", src,
                           "-o", str(tmp_path / "ts.o")])
