"""Timeline of ame_master_host per wave (bench batch): when each wave's H2D copy, kernels and D2H copy finish (B200)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from audio_mastering_engine_b200 import MasterPlan, synth, EQ_PRESETS
n_tr, fs, secs = 128, 48000, 180.0
n = int(secs * fs)
ids = list(range(n_tr))
mb = [t for t in ids if t % 2 == 1]; nb = [t for t in ids if t % 2 == 0]; ids = nb[:2] + mb + nb[2:]
settings = [synth.c4_settings(k, EQ_PRESETS) for k in ids]
dev = torch.device("cuda", 0)
tracks = synth.torch_track_batch(n_tr, secs, fs, dev)
h_in = torch.empty((n_tr * n, 2), dtype=torch.int16, pin_memory=True)
h_in.view(n_tr, n, 2).copy_(tracks)
del tracks
h_out = torch.empty_like(h_in, pin_memory=True)
cw = int(os.environ.get("AME_CHAIN_WARPS", "0"))
for waves in [int(a) for a in sys.argv[1:]] or [16]:
    plan = MasterPlan([n] * n_tr, fs, settings, host_io=True, n_waves=waves, chain_warps=cw)
    for _ in range(2): plan.master_host(h_in, h_out)
    t0 = time.perf_counter()
    for _ in range(3): plan.master_host(h_in, h_out)
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    plan.set_timing(True)
    plan.master_host(h_in, h_out)
    tl = plan.wave_timeline()
    kt, _ = plan.kernel_times()
    plan.set_timing(False)
    print(f"waves {waves} cw {cw}: host path {host_ms:.1f} ms (untimed)")
    print("  wave  h2d_done  run_start  run_done  d2h_done   (ms)")
    for w, r in enumerate(tl):
        print(f"  {w:4d}  {r[0]:8.1f}  {r[1]:8.1f}  {r[2]:8.1f}  {r[3]:8.1f}")
    print("  kernel sums:", {k: round(v[0], 1) for k, v in kt.items() if v[1]})
    plan.close()
