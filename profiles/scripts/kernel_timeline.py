"""Where every launch of one overlapped step sits on the device (ame_plan_kernel_timeline): per kernel, the time it
takes inside the overlapped step against the time it takes alone (one-slot plan), and how many launches are in flight."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, synth, EQ_PRESETS
from bench import batch_order, fill_batch

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=256)
ap.add_argument("--wave-tracks", type=int, default=32)
ap.add_argument("--slots", type=int, default=8)
ap.add_argument("--dump", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
fs, secs = 48000, 180.0
n = int(fs * secs)
ids = batch_order(list(range(a.tracks)), synth, EQ_PRESETS)
settings = [synth.c4_settings(t, EQ_PRESETS) for t in ids]
n_waves = (a.tracks + a.wave_tracks - 1) // a.wave_tracks
plan = MasterPlan([n] * a.tracks, fs, settings, n_waves=n_waves, n_slots=a.slots)
d_in = torch.zeros((plan.total_frames, 2), dtype=torch.int16, device=dev)
fill_batch(torch, synth, d_in, ids, n, secs, fs, dev)
d_out = torch.empty_like(d_in)
for _ in range(3):
    plan.master_device(d_in, d_out, fetch_results=False)
torch.cuda.synchronize()
plan.set_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    plan.master_device(d_in, d_out, fetch_results=False)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / 2
tl = plan.kernel_timeline(1)
names = sorted({k for w in tl for k in w})
t_end = max(v[1] for w in tl for v in w.values())
print(f"tracks {a.tracks} waves {plan.n_waves} slots {plan.n_slots}: {step_ms:.2f} ms per step (timeline span {t_end:.2f} ms)")
print(f"{'kernel':20s} {'launches':>8s} {'sum ms':>9s} {'mean ms':>8s} {'min':>7s} {'max':>7s}")
for k in names:
    d = np.array([w[k][1] - w[k][0] for w in tl if k in w])
    print(f"{k:20s} {len(d):8d} {d.sum():9.2f} {d.mean():8.3f} {d.min():7.3f} {d.max():7.3f}")
# launches in flight over time, per kernel
grid = np.linspace(0, t_end, 2001)[:-1]
tot = np.zeros_like(grid)
print("\nmean launches in flight:")
for k in names:
    c = np.zeros_like(grid)
    for w in tl:
        if k in w:
            c += (grid >= w[k][0]) & (grid < w[k][1])
    tot += c
    print(f"  {k:20s} {c.mean():6.2f}")
print(f"  {'all':20s} {tot.mean():6.2f}   (fraction of the step with nothing in flight: {(tot == 0).mean():.3f})")
print("\nper wave: begin of first kernel, end of last, busy share")
for i, w in enumerate(tl):
    b = min(v[0] for v in w.values()); e = max(v[1] for v in w.values())
    busy = sum(v[1] - v[0] for v in w.values())
    print(f"  wave {i:3d} slot {i % plan.n_slots}: {b:8.2f} -> {e:8.2f}  ({busy / max(e - b, 1e-9):.2f})")
if a.dump:
    json.dump(tl, open(a.dump, "w"))
plan.close()
