"""C3 split by time into 8 shards on ONE GPU (sharding.master_time_sharded_local) against the single plan: where do the
bytes differ, if anywhere?  (Diagnosis of a bit_identical_to_single_plan = false in bench.py --gpus 8.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_mastering_engine_b200 import MasterPlan, synth, sharding

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3600.0
fs = 96000
dev = torch.device("cuda", 0)
s = synth.c2_settings()
track = synth.torch_track_batch(1, secs, fs, dev, first_track_id=7)[0]
plan = MasterPlan([track.shape[0]], fs, s, device=0)
d_out = torch.empty_like(track)
info1 = plan.master_device(track, d_out)[0]
d_pre1 = torch.empty_like(track)
plan.close()
one = d_out.cpu().numpy()
del d_out
x = track.cpu().numpy()
del track
torch.cuda.empty_cache()
many, infon = sharding.master_time_sharded_local(x, fs, s, world)
diff = np.nonzero(np.any(one != many, axis=1))[0]
print(f"world {world}: {len(diff)} frames differ; input_i {info1['input_i']!r} vs {infon['input_i']!r}")
spans = sharding.plan_time_shards(len(x), fs, world, 30)
for i in diff[:20]:
    r = next(k for k, (lo, hi) in enumerate(spans) if lo <= i < hi)
    print(f"  frame {i} (shard {r}, {i - spans[r][0]} into it, chunk offset {i % (30 * fs)}): single {one[i]} sharded {many[i]}")
if len(diff):
    d = diff
    print("  runs:", np.split(d, np.nonzero(np.diff(d) > 1)[0] + 1)[:5])
