"""The same track (config C3) through the per-stage entry points with two different tilings of the time-parallel filters:
which stage's output differs, where, and by how much?  (Results must not depend on the tiling.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, synth

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
fs = 96000
dev = torch.device("cuda", 0)
s = synth.c2_settings()
track = synth.torch_track_batch(1, secs, fs, dev, first_track_id=7)[0]
n = track.shape[0]

def run(**opts):
    plan = MasterPlan([n], fs, s, device=0, **opts)
    d_pre = torch.zeros((plan.total_frames, 2), dtype=torch.int16, device=dev)
    d_bands = torch.zeros((3, plan.mb_frames, 2), dtype=torch.int16, device=dev)
    plan.stage_eq(track, d_pre)
    torch.cuda.synchronize()
    eq = d_pre[:n].clone()
    plan.stage_band_split(d_pre, d_bands)
    torch.cuda.synchronize()
    bands = d_bands[:, :n].clone()
    plan.stage_compress(d_bands, d_pre)
    torch.cuda.synchronize()
    out = d_pre[:n].clone()
    plan.close()
    return eq, bands, out

a = run()
for name, opts in (("eq tiles 2280", dict(eq_tile_frames=2280)), ("xover tiles 2280", dict(xover_tile_frames=2280)),
                   ("both 1144", dict(eq_tile_frames=1144, xover_tile_frames=1144)), ("both 40000", dict(eq_tile_frames=40000, xover_tile_frames=40000))):
    b = run(**opts)
    print(name)
    for what, x, y in (("  after EQ", a[0], b[0]), ("  low", a[1][0], b[1][0]), ("  mid", a[1][1], b[1][1]), ("  high", a[1][2], b[1][2]),
                       ("  after compressor", a[2], b[2])):
        d = torch.nonzero((x != y).any(dim=1)).flatten().cpu().numpy()
        msg = f"{what}: {len(d)} frames differ"
        for i in d[:4]:
            msg += f" | {i} (chunk offset {i % (30 * fs)}): {x[i].tolist()} vs {y[i].tolist()}"
        print(msg)
    del b
