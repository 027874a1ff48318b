"""Is the host path (ame_master_host) bound by PCIe or by the kernels?  For the bench batch and a host plan of W waves:
device-resident time of the SAME plan (master_device), its per-kernel sums, and the end-to-end time (B200)."""
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
d_in = tracks.view(n_tr * n, 2)
d_out = torch.empty_like(d_in)
h_out = torch.empty_like(h_in, pin_memory=True)
cw = int(os.environ.get("AME_CHAIN_WARPS", "0"))
for waves in [int(a) for a in sys.argv[1:]] or [8, 16, 32]:
    plan = MasterPlan([n] * n_tr, fs, settings, host_io=True, n_waves=waves, chain_warps=cw)
    for _ in range(2): plan.master_device(d_in, d_out, fetch_results=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): plan.master_device(d_in, d_out, fetch_results=False)
    e1.record(); torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 3
    plan.set_timing(True)
    plan.master_device(d_in, d_out, fetch_results=False)
    torch.cuda.synchronize()
    kt, _ = plan.kernel_times()
    plan.set_timing(False)
    plan.master_host(h_in, h_out)
    t0 = time.perf_counter()
    for _ in range(3): plan.master_host(h_in, h_out)
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    print(f"waves {waves:3d} cw {cw}: device-resident {dev_ms:7.1f} ms   host path {host_ms:7.1f} ms", flush=True)
    print("    kernel sums (ms, all waves):", {k: round(v[0], 1) for k, v in kt.items() if v[1]}, flush=True)
    plan.close()
