"""Generate tests/golden/*.npz by running the REFERENCE'S OWN functions (run in the build container only).

The upstream module imports ``pydub``, ``pydub.effects`` and ``ai_tagger`` at import time (none are
installed here), so they are stubbed in ``sys.modules``; its numpy/scipy functions
(audio_mastering_engine.py:250-309) then run unmodified on a minimal fake ``AudioSegment``.
``compress_dynamic_range`` is a pydub function, not reference code: the stub routes it to
``oracle.chain.compress_dynamic_range_py`` so the reference's own ``apply_multiband_compressor``
(crossover, int16 truncation, overlay order) is what produces the multiband fixtures.

Usage:  python tests/golden/make_golden.py [/root/reference]
The GPU box has no /root/reference: tests only read the committed .npz files.
"""
import array
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

with warnings.catch_warnings():
    warnings.simplefilter("ignore", DeprecationWarning)
    import audioop

from oracle import chain  # noqa: E402
from audio_mastering_engine_b200 import synth  # noqa: E402


class FakeSegment:
    """Just enough of pydub.AudioSegment for engine.py:250-309."""

    def __init__(self, data, frame_rate, channels=2, sample_width=2):
        self._data = bytes(data)
        self.frame_rate = frame_rate
        self.channels = channels
        self.sample_width = sample_width

    def get_array_of_samples(self):
        return array.array("h", self._data)

    def _spawn(self, data):
        return FakeSegment(data, self.frame_rate, self.channels, self.sample_width)

    def overlay(self, other):
        return self._spawn(audioop.add(self._data, other._data, self.sample_width))

    def pcm(self):
        return np.frombuffer(self._data, dtype=np.int16).reshape(-1, 2).copy()


def _stub_compress(seg, threshold=-20.0, ratio=4.0, attack=5.0, release=50.0):
    out = chain.compress_dynamic_range_py(seg.pcm(), seg.frame_rate, threshold, ratio, attack, release)
    return seg._spawn(out.tobytes())


def import_reference(ref_dir):
    pydub = types.ModuleType("pydub")
    pydub.AudioSegment = FakeSegment
    effects = types.ModuleType("pydub.effects")
    effects.compress_dynamic_range = _stub_compress
    sys.modules["pydub"] = pydub
    sys.modules["pydub.effects"] = effects
    sys.modules["ai_tagger"] = types.ModuleType("ai_tagger")
    sys.path.insert(0, ref_dir)
    import audio_mastering_engine as ref
    return ref


def seg(pcm, fs):
    return FakeSegment(np.ascontiguousarray(pcm, dtype=np.int16).tobytes(), fs)


def ref_process_chunk(ref, pcm, fs, settings):
    """The body of the reference's chunk loop (audio_mastering_engine.py:189-197), calling its functions."""
    chunk = seg(pcm, fs)
    taps = {}
    if settings.get("analog_character", 0) > 0:
        chunk = ref.apply_analog_character(chunk, settings.get("analog_character"))
    taps["warmth"] = chunk.pcm()
    chunk_samples = ref.audio_segment_to_float_array(chunk)
    processed = ref.apply_eq_to_samples(chunk_samples, chunk.frame_rate, settings)
    taps["eq"] = np.array(processed, copy=True)
    if settings.get("width", 1.0) != 1.0:
        processed = ref.apply_stereo_width(processed, settings.get("width"))
    out = ref.float_array_to_audio_segment(processed, chunk)
    taps["pre_multiband"] = out.pcm()
    if settings.get("multiband"):
        out = ref.apply_multiband_compressor(out, settings)
    taps["out"] = out.pcm()
    return taps


def main():
    ref_dir = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    ref = import_reference(ref_dir)
    assert ref.EQ_PRESETS == chain.EQ_PRESETS
    rng = np.random.default_rng(7)
    fixtures = {}

    # converters, incl. extremes (rows 2-3)
    edge = np.array([[-32768, 32767], [-32767, 32766], [0, 1], [-1, 1], [12345, -12345]], dtype=np.int16)
    pcm = np.concatenate([edge, rng.integers(-32768, 32768, size=(507, 2)).astype(np.int16)])
    f = ref.audio_segment_to_float_array(seg(pcm, 48000))
    fixtures["conv_in"] = pcm
    fixtures["conv_float"] = f
    ramp = np.linspace(-1.2, 1.2, 4001, dtype=np.float32).repeat(2).reshape(-1, 2)
    fixtures["conv_ramp_f32"] = ramp
    fixtures["conv_ramp_pcm_f32"] = ref.float_array_to_audio_segment(ramp, seg(pcm, 48000)).pcm()
    fixtures["conv_ramp_pcm_f64"] = ref.float_array_to_audio_segment(ramp.astype(np.float64) * 0.999, seg(pcm, 48000)).pcm()

    # per-stage and per-chunk cases
    cases = []
    eq_sets = {"flat": dict(bass_boost=0.0, mid_cut=0.0, presence_boost=0.0, treble_boost=0.0),
               "allboost": dict(synth.ALL_BOOST_EQ)}
    eq_sets.update({k: dict(v) for k, v in ref.EQ_PRESETS.items()})
    for fs, secs in ((44100, 0.20), (48000, 0.20), (96000, 0.12), (192000, 0.08)):
        x = synth.track(secs, fs, track_id=fs % 97, am_hz=20.0)
        x[:3] = [[32767, -32768], [-32768, 32767], [0, 0]]
        for name, eqs in eq_sets.items():
            if fs not in (44100, 48000) and name not in ("allboost", "Lo-Fi Haze"):
                continue
            for variant, extra in (("plain", dict(analog_character=0, width=1.0, multiband=False)),
                                   ("full", dict(analog_character=25, width=1.2, multiband=True,
                                                 **synth.DEFAULT_MULTIBAND))):
                if variant == "full" and name not in ("allboost", "Vocal Clarity", "Lo-Fi Haze"):
                    continue
                settings = dict(eqs, **extra)
                cases.append((fs, name, variant, x, settings))
    # warmth 100 %, width extremes
    x48 = synth.track(0.1, 48000, track_id=5)
    cases.append((48000, "flat", "warm100", x48, dict(eq_sets["flat"], analog_character=100, width=1.0, multiband=False)))
    cases.append((48000, "flat", "width0", x48, dict(eq_sets["flat"], analog_character=0, width=0.0, multiband=False)))
    cases.append((48000, "flat", "width2", (x48.astype(np.int32) * 3).clip(-32768, 32767).astype(np.int16),
                  dict(eq_sets["flat"], analog_character=0, width=2.0, multiband=False)))
    # compressor extremes on a hot signal (C5 settings)
    hot = (synth.track(0.15, 48000, track_id=9, am_hz=15.0, am_db=12.0).astype(np.int32) * 3).clip(-32768, 32767).astype(np.int16)
    hot[3000:5000] = 0
    for nm, th, ra in (("comp_hard", -40.0, 10.0), ("comp_unity", 0.0, 1.0)):
        s = dict(eq_sets["flat"], analog_character=0, width=1.0, multiband=True, low_thresh=th, low_ratio=ra,
                 mid_thresh=th, mid_ratio=ra, high_thresh=th, high_ratio=ra)
        cases.append((48000, "flat", nm, hot, s))

    index = []
    for k, (fs, name, variant, x, settings) in enumerate(cases):
        taps = ref_process_chunk(ref, x, fs, settings)
        key = f"case{k:02d}"
        fixtures[key + "_in"] = x
        for tname, v in taps.items():
            fixtures[f"{key}_{tname}"] = v
        index.append(dict(key=key, fs=fs, eq=name, variant=variant, settings=settings))
        print(key, fs, name, variant, "out crc", int(np.abs(taps["out"].astype(np.int64)).sum()))

    np.savez_compressed(os.path.join(HERE, "reference_chain.npz"), **fixtures)
    import json
    with open(os.path.join(HERE, "reference_chain.json"), "w") as fh:
        json.dump(dict(reference="theouterlimitz/Audio-Mastering-Engine audio_mastering_engine.py:250-309",
                       numpy=np.__version__, scipy=__import__("scipy").__version__, cases=index), fh, indent=1)
    print("wrote", len(fixtures), "arrays")


if __name__ == "__main__":
    main()
