"""k_att_chain time of single 3 min tracks under the C4 settings sweep: queue kernel alone (chain_warps=-1) against the
speculative kernel + queue-kernel fallback (auto).  Some presets leave the mid band hovering at threshold, where the
attenuation never clamps again and the speculation cannot settle (B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, synth, EQ_PRESETS
fs, secs = 48000, 180.0
for tid in [int(a) for a in sys.argv[1:]] or [1, 3, 5, 7, 9, 11, 13, 15, 19, 37]:
    x = synth.track(secs, fs, tid, am_hz=2.0)
    s = synth.c4_settings(tid, EQ_PRESETS)
    row = []
    outs = []
    for cw in (-1, 0):
        plan = MasterPlan([len(x)], fs, s, chain_warps=cw)
        d_in = torch.from_numpy(plan.pack([x])).cuda(); d_out = torch.empty_like(d_in)
        for _ in range(2): plan.master_device(d_in, d_out, fetch_results=False)
        torch.cuda.synchronize()
        plan.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): plan.master_device(d_in, d_out, fetch_results=False)
        e1.record(); torch.cuda.synchronize()
        kt, _ = plan.kernel_times()
        row.append((e0.elapsed_time(e1) / 5, kt["k_att_chain"][0] / max(kt["k_att_chain"][1], 1)))
        outs.append(d_out.cpu().numpy().copy())
        plan.close()
    print(f"track {tid:3d}: queue {row[0][0]:6.2f} ms (chain {row[0][1]:6.2f})   spec+fallback {row[1][0]:6.2f} ms (chain {row[1][1]:6.2f})   same bytes {np.array_equal(outs[0], outs[1])}", flush=True)
