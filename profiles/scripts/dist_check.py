"""Run under torchrun on N GPUs: time-sharded mastering of one long track over NCCL must equal the
single-GPU result bit for bit.  Prints one OK line from rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
from audio_mastering_engine_b200 import master, synth, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fs, seconds = 96000, float(os.environ.get("DIST_SECONDS", "240"))
x = synth.track(seconds, fs, track_id=7, am_hz=2.0, drift_db=8.0, drift_period=90.0)       # C3-style, shortened
s = synth.c2_settings()
begin, out, info = sharding.master_time_sharded(x, fs, s)
ref, rinfo = master(x, fs, s, device=local)                                               # whole track on this GPU
same = np.array_equal(out, ref[begin:begin + len(out)])
flag = torch.tensor([int(same)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"dist_check world={world} fs={fs} seconds={seconds}: bit_identical={bool(flag.item())} "
          f"LUFS sharded {info['input_i']:.6f} single {rinfo['input_i']:.6f} blocks {info['n_blocks']}/{rinfo['n_blocks']}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
