"""k_att_chain time vs number of identical multiband tracks (contention check), 30 s chunks at 48 kHz (B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, synth
fs = 48000; n = 30 * fs
x = synth.track(30.0, fs, 1, am_hz=2.0)
s = dict(bass_boost=-1.0, mid_cut=2.0, presence_boost=2.5, treble_boost=1.0, analog_character=0, width=0.8, lufs=-14.0,
         multiband=True, **synth.DEFAULT_MULTIBAND)
for copies in (1, 16, 64, 192, 384):
    plan = MasterPlan([n] * copies, fs, s)
    d_in = torch.from_numpy(plan.pack([x] * copies)).cuda(); d_out = torch.empty_like(d_in)
    for _ in range(2): plan.master_device(d_in, d_out, fetch_results=False)
    plan.set_timing(True)
    for _ in range(3): plan.master_device(d_in, d_out, fetch_results=False)
    kt, _ = plan.kernel_times()
    ms = kt["k_att_chain"][0] / kt["k_att_chain"][1]
    print(f"{copies:4d} chunks ({copies*3} chains): k_att_chain {ms:7.3f} ms = {ms*1e-3*1.965e9/n:5.1f} cycles/frame", flush=True)
    plan.close()
