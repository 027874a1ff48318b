"""e2e (host buffers) throughput of the 128-track batch vs number of plan waves (B200)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from audio_mastering_engine_b200 import MasterPlan, synth, EQ_PRESETS
n_tr, fs, secs = 128, 48000, 180.0
n = int(secs * fs)
ids = list(range(n_tr))
if os.environ.get('ORDERED') == '1':
    mb = [t for t in ids if t % 2 == 1]; nb = [t for t in ids if t % 2 == 0]; ids = nb[:2] + mb + nb[2:]
settings = [synth.c4_settings(k, EQ_PRESETS) for k in ids]
dev = torch.device("cuda", 0)
tracks = synth.torch_track_batch(n_tr, secs, fs, dev)
h_in = torch.empty((n_tr * n, 2), dtype=torch.int16, pin_memory=True)
h_in.view(n_tr, n, 2).copy_(tracks)
del tracks
h_out = torch.empty_like(h_in, pin_memory=True)
for waves in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16]:
    plan = MasterPlan([n] * n_tr, fs, settings, host_io=True, n_waves=waves)
    plan.master_host(h_in, h_out)
    t0 = time.perf_counter()
    for _ in range(3):
        plan.master_host(h_in, h_out)
    dt = (time.perf_counter() - t0) / 3
    print(f"waves {waves:3d}: {dt*1e3:7.1f} ms/step  {n_tr*secs/dt:9.0f} x realtime", flush=True)
    plan.close()
