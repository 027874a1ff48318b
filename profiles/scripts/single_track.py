"""Per-kernel times for ONE track (BASELINE configs C1 / C2 / C5), device-resident and via the host API (B200)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, synth
cases = {"c1": (44100, 30.0, synth.c1_settings(), None), "c2": (48000, 180.0, synth.c2_settings(), 2.0),
         "c2lim": (48000, 180.0, dict(synth.c2_settings(), limiter=True, true_peak=True), 2.0),
         "c2hot": (48000, 180.0, dict(synth.c2_settings(), lufs=-6.0, limiter=True, true_peak=True), 2.0),   # clips: the limiter's worst case
         "c5": (192000, 600.0, dict(synth.ALL_BOOST_EQ, analog_character=25, width=1.2, lufs=-14.0, multiband=True,
                                    low_thresh=-40.0, low_ratio=10.0, mid_thresh=-40.0, mid_ratio=10.0, high_thresh=-40.0, high_ratio=10.0), "stress")}
cases["c3"] = (96000, 3600.0, synth.c2_settings(), "device")         # generated on the GPU (the numpy recipe takes minutes at this length)
cw = int(os.environ.get("AME_CHAIN_WARPS", "0"))
for name in sys.argv[1:] or ["c1", "c2"]:
    fs, secs, s, am = cases[name]
    if am == "device":
        x = synth.torch_track_batch(1, secs, fs, torch.device("cuda", 0))[0].cpu().numpy()
        torch.cuda.empty_cache()
    else:
        x = synth.stress_track(secs, fs, 0) if am == "stress" else synth.track(secs, fs, 0, am_hz=am)
    plan = MasterPlan([len(x)], fs, s, host_io=True, chain_warps=cw)
    h_in = torch.from_numpy(plan.pack([x])).pin_memory(); h_out = torch.empty_like(h_in).pin_memory()
    d_in = h_in.cuda(); d_out = torch.empty_like(d_in)
    for _ in range(3): plan.master_device(d_in, d_out, fetch_results=False)
    torch.cuda.synchronize()
    plan.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): plan.master_device(d_in, d_out, fetch_results=False)
    e1.record(); torch.cuda.synchronize()
    kt, _ = plan.kernel_times()
    plan.set_timing(False)
    dev_ms = e0.elapsed_time(e1) / 10
    for _ in range(2): plan.master_host(h_in, h_out)
    t0 = time.perf_counter()
    for _ in range(10): plan.master_host(h_in, h_out)
    host_ms = (time.perf_counter() - t0) / 10 * 1e3
    print(f"{name}: device-resident {dev_ms:.3f} ms ({secs/dev_ms*1e3:.0f} x realtime), host API pinned {host_ms:.3f} ms ({secs/host_ms*1e3:.0f} x realtime)")
    print("   ", {k: round(v[0] / max(v[1], 1), 3) for k, v in kt.items() if v[1]})
    plan.close()
