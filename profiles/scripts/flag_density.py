"""Fraction of frames above the compressor threshold per (track, band) of the bench batch (B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from audio_mastering_engine_b200 import MasterPlan, synth, EQ_PRESETS
n_tr, fs, secs = 24, 48000, 180.0
n = int(secs * fs)
settings = [synth.c4_settings(k, EQ_PRESETS) for k in range(n_tr)]
dev = torch.device("cuda", 0)
d_in = synth.torch_track_batch(n_tr, secs, fs, dev).view(n_tr * n, 2).contiguous()
plan = MasterPlan([n] * n_tr, fs, settings)
d_pre = torch.zeros_like(d_in)
d_bands = torch.zeros((3, plan.mb_frames, 2), dtype=torch.int16, device=dev)
plan.stage_eq(d_in, d_pre); plan.stage_band_split(d_pre, d_bands); torch.cuda.synchronize()
thr = [32768 * 10 ** (t / 20) for t in (-25.0, -20.0, -15.0)]
for t in range(n_tr):
    off = plan.mb_offset(t)
    if off < 0: continue
    row = []
    for b in range(3):
        x = d_bands[b, off:off + n].to(torch.float64)
        e = (x * x).sum(1)
        c = torch.cumsum(e, 0)
        s = c.clone(); s[240:] = c[240:] - c[:-240]
        s = torch.roll(s, 1); s[0] = 0
        frac = float(((s / 480.0) >= (np.floor(thr[b]) + 1) ** 2).double().mean())
        row.append(round(frac, 3))
    print(t, {k: settings[t][k] for k in ("bass_boost", "analog_character")}, row)
plan.close()
