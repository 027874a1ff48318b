#!/usr/bin/env python
"""Benchmark of the mastering DSP chain (BASELINE.json metric: audio-seconds mastered per wall-second).

  python bench.py [--gpus N] [--steps K] [--warmup W]                      # this build, N GPUs of one node
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                                      # the reference's CPU path

Workload (config.workload): BASELINE config C4 - a batch of synthetic 3 min stereo 48 kHz tracks with
the C4 settings sweep, sharded BY TRACK: 128 tracks per GPU (1024 tracks at 8 GPUs), no data-path
collective, weak scaling.  A step = one pass of the whole chain (EQ + warmth + width + multiband
compressor + BS.1770 loudness + gain) over the rank's batch.
  value : device-resident (inputs already in HBM), CUDA-event timed, max over ranks.
  e2e   : same batch through the public host API (ame_master_host via MasterPlan.master_host) with pinned
          HOST buffers: H2D + chain + D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio-sec mastered/sec (x realtime)"
UNIT = "x realtime"
B_ALG_CHAIN = 16       # bytes per stereo frame, whole chain with normalisation
# DRAM bytes per launch measured once with `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum) on the
# 128-track workload with ONE plan wave; scaled by 1 / waves below.  Source: profiles/r01e_summary.md.
KERNEL_NCU_TRAFFIC_1WAVE = {"k_eq": 9.242e9, "k_band_split": 9.343e9, "k_window_flag": 9.941e9, "k_att_chain": 6.591e9,
                            "k_compress_apply": 13.273e9, "k_kweight_energy": 6.363e9}
KERNEL_ALG_BYTES = {   # per-kernel algorithmic bytes per frame it processes
    "k_eq": 8, "k_band_split": 16, "k_window_flag": 18, "k_att_chain": 6, "k_compress_apply": 22,
    "k_kweight_energy": 4, "k_apply_gain": 8}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks-per-gpu", type=int, default=128)
    ap.add_argument("--seconds", type=float, default=180.0)
    ap.add_argument("--fs", type=int, default=48000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-waves", type=int, default=32)
    ap.add_argument("--chain-warps", type=int, default=0, help="ame_plan_options.chain_warps (0 = auto, -1 = queue kernel)")
    ap.add_argument("--waves", type=int, default=6, help="plan waves of the device-resident path")
    ap.add_argument("--slots", type=int, default=0, help="workspace slots (0 = min(waves, 4))")
    ap.add_argument("--cpu-sample-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's NUMA node")
    ap.add_argument("--kw-tile", type=int, default=0, help="ame_plan_options.kw_tile_subblocks (0 = auto)")
    ap.add_argument("--eq-tile", type=int, default=0, help="ame_plan_options.eq_tile_frames (0 = auto)")
    ap.add_argument("--xover-tile", type=int, default=0, help="ame_plan_options.xover_tile_frames (0 = auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port: numpy/scipy + pydub's Python loop on
    audioop), all host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import cpu_arm
    cores = os.cpu_count() or 1
    pool = cpu_arm.make_pool(cores)
    ids = list(range(cores))
    secs, fs = args.cpu_sample_seconds, args.fs
    for _ in range(max(args.warmup, 1) if pool is not None else 0):
        cpu_arm.run_step(pool, ids[:cores], min(secs, 1.0), fs)      # warm the workers (imports, filter design)
    audio = wall = 0.0
    for _ in range(args.steps):
        a, w = cpu_arm.run_step(pool, ids, secs, fs)
        audio += a
        wall += w
    if pool is not None:
        pool.shutdown()
    value = audio / wall
    sample = (f"{cores} tracks x {secs:g} s of the C4 sweep per step, one per core, {args.steps} steps; "
              "ffmpeg split/concat/loudnorm substituted by numpy restatements (binary absent)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: batch of 3 min stereo 48 kHz tracks, settings sweep (bounded CPU sample)",
                       "fs": fs, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from audio_mastering_engine_b200 import MasterPlan, synth, EQ_PRESETS, bind_host_to_gpu_numa

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = None if args.no_numa_bind else bind_host_to_gpu_numa(local_rank)   # before any pinned allocation
    if world > 1:
        # NCCL writes its version banner to stdout when the communicator is created: send fd 1 to stderr until the
        # first collective is through, so that stdout carries nothing but the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_tr, fs, secs = args.tracks_per_gpu, args.fs, args.seconds
    n = int(round(secs * fs))
    first = rank * n_tr
    # the C4 sweep of this rank, batched by cost as a host batcher would: two cheap tracks first (the host path can
    # start copying results back early), then the multiband tracks (their sequential compressor kernel runs under
    # the later copies), then the remaining tracks (short drain after the last copy-in)
    ids = list(range(first, first + n_tr))
    mb = [t for t in ids if synth.c4_settings(t, EQ_PRESETS)["multiband"]]
    nb = [t for t in ids if not synth.c4_settings(t, EQ_PRESETS)["multiband"]]
    ids = nb[:2] + mb + nb[2:]
    settings = [synth.c4_settings(t, EQ_PRESETS) for t in ids]
    plan = MasterPlan([n] * n_tr, fs, settings, device=local_rank, n_waves=args.waves, chain_warps=args.chain_warps,
                      kw_tile_subblocks=args.kw_tile, eq_tile_frames=args.eq_tile, xover_tile_frames=args.xover_tile,
                      n_slots=args.slots)
    assert plan.total_frames == n_tr * ((n + 7) // 8 * 8)
    tracks = synth.torch_track_batch(n_tr, secs, fs, dev, first_track_id=first)        # [n_tr, n, 2] int16
    d_in = torch.zeros((plan.total_frames, 2), dtype=torch.int16, device=dev)
    d_in.view(n_tr, -1, 2)[:, :n] = tracks
    del tracks
    d_out = torch.empty_like(d_in)
    stream = torch.cuda.current_stream().cuda_stream
    frames_rank = n_tr * n
    audio_rank = n_tr * secs

    for _ in range(max(args.warmup, 3)):
        plan.master_device(d_in, d_out, stream=stream, fetch_results=False)
    barrier()
    plan.set_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.master_device(d_in, d_out, stream=stream, fetch_results=False)
    e1.record()
    barrier()
    clocks = sampler.finish()
    ms = e0.elapsed_time(e1)
    ktimes, ksteps = plan.kernel_times()
    plan.set_timing(False)
    launches = plan.launch_count * args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * audio_rank * args.steps / (ms_max * 1e-3)

    # ---- end to end through the host API: pinned host buffers, H2D + chain + D2H per step ----------
    e2e = None
    if not args.no_e2e:
        # the host API on its own plan: same batch, split into waves so copies and kernels overlap
        hplan = MasterPlan([n] * n_tr, fs, settings, device=local_rank, host_io=True, n_waves=args.e2e_waves, chain_warps=args.chain_warps)
        h_in = torch.empty((plan.total_frames, 2), dtype=torch.int16, pin_memory=True)
        h_out = torch.empty_like(h_in, pin_memory=True)
        h_in.copy_(d_in)
        torch.cuda.synchronize()
        hplan.master_host(h_in, h_out)                                   # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = hplan.master_host(h_in, h_out)                         # synchronous: returns after D2H
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(h_out.to(dev), d_out))
        e2e = {"value": world * audio_rank * args.e2e_steps / float(tw.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h_in.numel() * 2), "d2h_bytes_per_step": int(h_out.numel() * 2 + 48 * n_tr),
               "steps": args.e2e_steps, "waves": args.e2e_waves, "numa_node": numa_node, "matches_device_path": same,
               "first_track_lufs": res[0]["input_i"]}
        hplan.close()
        del h_in, h_out

    # ---- roofline of the dominant kernel (live CUDA-event durations from the timed region) ---------
    peak, peak_src = measured_peak_gbs()
    mb_frames = sum(1 for s in settings if s["multiband"]) * n
    per_kernel = {}
    for name, (msum, cnt) in ktimes.items():
        if cnt:
            per_kernel[name] = msum / cnt
    dom = max(per_kernel, key=per_kernel.get)
    # one launch of a kernel covers one plan wave (1 / waves of the rank's frames)
    dom_frames = (frames_rank if dom in ("k_eq", "k_kweight_energy", "k_apply_gain") else mb_frames) / max(args.waves, 1)
    alg_bytes = KERNEL_ALG_BYTES.get(dom, 8) * dom_frames
    achieved = alg_bytes / (per_kernel[dom] * 1e-3) / 1e9
    chain_gbs = B_ALG_CHAIN * frames_rank * args.steps / (ms * 1e-3) / 1e9
    default_shape = (n_tr == 128 and secs == 180.0 and fs == 48000)
    traffic = KERNEL_NCU_TRAFFIC_1WAVE.get(dom) if default_shape else None
    per_kernel_roof = {}
    for name, kms in per_kernel.items():
        if name in KERNEL_ALG_BYTES:
            fr = (frames_rank if name in ("k_eq", "k_kweight_energy", "k_apply_gain") else mb_frames) / max(args.waves, 1)
            gbs = KERNEL_ALG_BYTES[name] * fr / (kms * 1e-3) / 1e9
            per_kernel_roof[name] = {"achieved": round(gbs, 1), "frac": round(gbs / peak, 4)}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic / max(args.waves, 1)) if traffic else None,
                "traffic_source": "ncu --set full on this workload, 1 plan wave, divided by waves (profiles/r01e_summary.md)" if traffic else None,
                "peak_source": peak_src, "kernel_ms": per_kernel[dom], "per_kernel": per_kernel_roof,
                "algorithmic_bytes_per_launch": alg_bytes,
                "chain": {"achieved": chain_gbs, "frac": chain_gbs / peak, "bytes_per_frame": B_ALG_CHAIN},
                "plan_waves": args.waves,
                "note": "kernel_ms = average duration of ONE launch (one plan wave) from CUDA events on that wave's stream "
                        "inside the timed region; launches of different waves overlap, so shares add up to more than 1",
                "kernel_ms_all": {k: round(v, 4) for k, v in per_kernel.items()},
                "kernel_share_of_step": {k: round(v * args.waves / (ms / args.steps), 4) for k, v in per_kernel.items()}}

    # ---- CPU baseline (rank 0, N=1 only): single thread, as the reference runs ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_arm
        ids = [0, 1, 2, 3]
        a, w = cpu_arm.run_step(None, ids, 5.0, fs)
        cpu = {"value": a / w, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "tracks 0-3 of the C4 sweep x 5 s each, sequential on one core (2 of 4 multiband); "
                         "oracle port = reference numpy/scipy stages + pydub Python loop on audioop; ffmpeg stages "
                         "substituted by numpy restatements (binary absent)",
               "host_cpus": os.cpu_count()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C4 shard: {n_tr} synthetic {secs:g} s stereo {fs} Hz tracks per GPU "
                                       f"({n_tr * world} tracks total), C4 settings sweep, 30 s chunks, sharded by track",
                           "tracks_per_gpu": n_tr, "seconds": secs, "fs": fs, "chunk_seconds": 30,
                           "parallelism": f"by-track x{world}, no collective", "plan_waves": args.waves,
                           "l2": f"inputs larger than L2 ({d_in.numel() * 2 / 1e9:.2f} GB per GPU per pass)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "workspace_gb": round(plan.workspace_bytes / 1e9, 2), "chain_stats": plan.chain_stats()}
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
