"""Small mixed batch through the whole path (for compute-sanitizer memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from audio_mastering_engine_b200 import master, synth, EQ_PRESETS, sharding
tracks = [synth.track(0.6 + 0.21 * k, fs, track_id=k, am_hz=3.0)[: int((0.6 + 0.21 * k) * fs) - k] for k, fs in enumerate((44100, 48000, 22050, 96000, 48000))]
fss = [44100, 48000, 22050, 96000, 48000]
sets = [synth.c4_settings(k + 1, EQ_PRESETS) for k in range(5)]
sets[2]["analog_character"] = 0
outs, infos = master(tracks, fss, sets, chunk_seconds=0.25, n_waves=2)
print("batch ok", [round(i["input_i"], 3) for i in infos])
x = synth.track(3.0, 48000, 9, am_hz=2.0)
out, info = sharding.master_time_sharded_local(x, 48000, synth.c2_settings(), 3, chunk_seconds=0.5)
print("shards ok", round(info["input_i"], 3), out.shape)
