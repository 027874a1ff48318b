"""Device-resident throughput of the 128-track batch vs number of plan waves (B200)."""
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
ref = None
big = os.environ.get('BIG_TILES') == '1'
for waves in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    kw = dict(eq_tile_frames=29440, xover_tile_frames=9800) if big else {}
    plan = MasterPlan([n] * n_tr, fs, settings, n_waves=waves, **kw)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.master_device(d_in, d_out, stream=st, fetch_results=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.master_device(d_in, d_out, stream=st, fetch_results=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    same = True if ref is None else bool(torch.equal(ref, d_out))
    if ref is None: ref = d_out.clone()
    print(f"waves {waves:3d}: {ms:7.2f} ms/step  {n_tr*secs/ms*1e3:9.0f} x realtime  same_as_1_wave={same}", flush=True)
    plan.close()
