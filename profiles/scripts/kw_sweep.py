"""k_kweight_energy / k_band_split time vs tile size on the 128-track batch (B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from audio_mastering_engine_b200 import MasterPlan, synth, EQ_PRESETS
n_tr, fs, secs = 128, 48000, 180.0
n = int(secs * fs)
settings = [synth.c4_settings(k, EQ_PRESETS) for k in range(n_tr)]
dev = torch.device("cuda", 0)
d_in = synth.torch_track_batch(n_tr, secs, fs, dev).view(n_tr * n, 2).contiguous()
d_out = torch.empty_like(d_in)
for kw in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8, 12]:
    plan = MasterPlan([n] * n_tr, fs, settings, kw_tile_subblocks=kw)
    for _ in range(2): plan.master_device(d_in, d_out, fetch_results=False)
    plan.set_timing(True)
    for _ in range(3): plan.master_device(d_in, d_out, fetch_results=False)
    kt, _ = plan.kernel_times()
    print(f"kw tile {kw:3d} sub-blocks: k_kweight_energy {kt['k_kweight_energy'][0]/3:.3f} ms   k_band_split {kt['k_band_split'][0]/3:.3f} ms  k_eq {kt['k_eq'][0]/3:.3f}", flush=True)
    plan.close()
