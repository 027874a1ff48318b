"""Per-kernel times of the compressor path on single 30 s chunks with controlled flag density (B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, synth

fs = 48000
n = 30 * fs
rng = np.random.default_rng(0)
base = synth.track(30.0, fs, 1)
for name, gain_db in (("quiet (no flags)", -30.0), ("nominal", 0.0), ("loud (dense flags)", 18.0)):
    x = np.clip(base.astype(np.float64) * 10 ** (gain_db / 20), -32768, 32767).astype(np.int16)
    s = dict(bass_boost=0.0, mid_cut=0.0, presence_boost=0.0, treble_boost=0.0, analog_character=0, width=1.0, lufs=-14.0,
             multiband=True, **synth.DEFAULT_MULTIBAND)
    plan = MasterPlan([n], fs, s, device=0)
    d_in = torch.from_numpy(plan.pack([x])).cuda()
    d_out = torch.empty_like(d_in)
    for _ in range(3):
        plan.master_device(d_in, d_out, fetch_results=False)
    plan.set_timing(True)
    for _ in range(5):
        plan.master_device(d_in, d_out, fetch_results=False)
    kt, steps = plan.kernel_times()
    ms = {k: v[0] / max(v[1], 1) for k, v in kt.items()}
    print(name, {k: round(v, 3) for k, v in ms.items() if v > 0.005})
    c = ms["k_att_chain"] * 1e-3 * 1.965e9
    print("   k_att_chain cycles per frame (slowest of 3 band chains): %.1f" % (c / n))
    plan.close()
