import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, synth
fs, secs = 96000, 3600.0
dev = torch.device("cuda", 0)
s = synth.c2_settings()
track = synth.torch_track_batch(1, secs, fs, dev, first_track_id=7)[0]
n = track.shape[0]
plan = MasterPlan([n], fs, s, device=0)
d_pre = torch.zeros((plan.total_frames, 2), dtype=torch.int16, device=dev)
plan.stage_eq(track, d_pre)
torch.cuda.synchronize()
cf = 30 * fs
for c in (14, 53):
    np.save(f"gpurun_out/c3_eq_chunk{c}.npy", d_pre[c * cf:(c + 1) * cf].cpu().numpy())
plan.close()
