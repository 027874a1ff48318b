"""Wall time of building a plan (filter design, job tables, workspace allocation) against running it, one track (B200)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, master, synth
torch.zeros(1).cuda()
for name, fs, secs, s, am in (("c1", 44100, 30.0, synth.c1_settings(), None), ("c2", 48000, 180.0, synth.c2_settings(), 2.0)):
    x = synth.track(secs, fs, 0, am_hz=am)
    for rep in range(3):
        t0 = time.perf_counter()
        plan = MasterPlan([len(x)], fs, s, host_io=True)
        t1 = time.perf_counter()
        h_in = plan.pack([x]); h_out = np.empty_like(h_in)
        t2 = time.perf_counter()
        plan.master_host(h_in, h_out)
        t3 = time.perf_counter()
        plan.close()
        t4 = time.perf_counter()
        print(f"{name} rep {rep}: plan build {1e3*(t1-t0):7.1f} ms  pack {1e3*(t2-t1):6.1f}  master_host (pageable) {1e3*(t3-t2):6.1f}  close {1e3*(t4-t3):6.1f}", flush=True)
    t0 = time.perf_counter(); out, info = master(x, fs, s); t1 = time.perf_counter()
    print(f"{name}: master() end to end {1e3*(t1-t0):.1f} ms for {secs:.0f} s of audio", flush=True)
