"""Null test of the CUDA path against the CPU oracle at the FULL sizes of BASELINE.json configs C1, C2, C3, C5
(C3 as 8 time shards emulated on one GPU; C4 is 1024 x C2).  Oracle chunks run on all host cores.
Usage: python profiles/scripts/full_config_parity.py [c1 c2 c3 c5]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from concurrent.futures import ProcessPoolExecutor
import multiprocessing as mp


def _chunk(args):
    x, fs, settings = args
    from oracle import chain
    return chain.process_chunk(x, fs, settings)


def oracle_master(x, fs, settings, pool):
    from oracle import chain
    cf = 30 * fs
    parts = list(pool.map(_chunk, [(x[s:s + cf], fs, settings) for s in range(0, len(x), cf)]))
    pre = np.concatenate(parts)
    info = {}
    out = chain.normalize(pre, fs, settings["lufs"], info) if settings.get("lufs") is not None else pre
    return out, info


def report(name, out, info, ref, rinfo, t_gpu, t_cpu, seconds):
    d = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    mx = int(d.max())
    db = 20 * np.log10(max(mx, 1e-9) / 32768.0) if mx else float("-inf")
    print(f"{name}: frames {len(out)}  max|diff| {mx} LSB ({db:.1f} dBFS)  differing samples {float((d > 0).mean()):.2e}  "
          f"LUFS gpu {info['input_i']:.6f} oracle {rinfo['input_i']:.6f} (delta {abs(info['input_i'] - rinfo['input_i']):.2e})  "
          f"gain {info['gain']:.9f}/{rinfo['gain']:.9f}  gpu {t_gpu*1e3:.1f} ms ({seconds / t_gpu:.0f} x realtime incl. H2D/D2H, plan build excluded)  "
          f"oracle {t_cpu:.1f} s", flush=True)
    assert mx <= 3 and abs(info["input_i"] - rinfo["input_i"]) <= 0.01


def gpu_master(x, fs, s, **kw):
    from audio_mastering_engine_b200 import MasterPlan
    plan = MasterPlan([len(x)], fs, s, host_io=True, **kw)
    h_in = plan.pack([x]); h_out = np.empty_like(h_in)
    plan.master_host(h_in, h_out)
    t0 = time.perf_counter()
    info = plan.master_host(h_in, h_out)[0]
    dt = time.perf_counter() - t0
    out = plan.unpack(h_out)[0]
    plan.close()
    return out, info, dt


def main():
    from audio_mastering_engine_b200 import synth, sharding
    which = sys.argv[1:] or ["c1", "c2", "c3", "c5"]
    with ProcessPoolExecutor(max_workers=os.cpu_count(), mp_context=mp.get_context("spawn")) as pool:
        if "c1" in which:
            fs, secs = 44100, 30.0
            x, s = synth.track(secs, fs, 0), synth.c1_settings()
            out, info, dt = gpu_master(x, fs, s)
            t0 = time.time(); ref, rinfo = oracle_master(x, fs, s, pool); tc = time.time() - t0
            report("C1 30 s 44.1 kHz default chain", out, info, ref, rinfo, dt, tc, secs)
        if "c2" in which:
            fs, secs = 48000, 180.0
            x, s = synth.track(secs, fs, 0, am_hz=2.0), synth.c2_settings()
            out, info, dt = gpu_master(x, fs, s)
            t0 = time.time(); ref, rinfo = oracle_master(x, fs, s, pool); tc = time.time() - t0
            report("C2 3 min 48 kHz multiband", out, info, ref, rinfo, dt, tc, secs)
        if "c5" in which:
            fs, secs = 192000, 600.0
            x = synth.stress_track(secs, fs, 3)
            for th, ra in ((-40.0, 10.0), (0.0, 1.0)):
                s = dict(synth.ALL_BOOST_EQ, analog_character=25, width=1.2, lufs=-14.0, multiband=True, low_thresh=th, low_ratio=ra,
                         mid_thresh=th, mid_ratio=ra, high_thresh=th, high_ratio=ra)
                out, info, dt = gpu_master(x, fs, s)
                t0 = time.time(); ref, rinfo = oracle_master(x, fs, s, pool); tc = time.time() - t0
                report(f"C5 10 min 192 kHz stress thresh {th} ratio {ra}", out, info, ref, rinfo, dt, tc, secs)
        if "c3" in which:
            fs, secs = 96000, 3600.0
            x = synth.track(secs, fs, 2, am_hz=2.0, drift_db=8.0, drift_period=90.0)
            s = synth.c2_settings()
            import torch
            t0 = time.perf_counter()
            out, info = sharding.master_time_sharded_local(x, fs, s, 8)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            t0 = time.time(); ref, rinfo = oracle_master(x, fs, s, pool); tc = time.time() - t0
            report("C3 60 min 96 kHz, 8 time shards (emulated on one GPU, time incl. plan builds)", out, info, ref, rinfo, dt, tc, secs)


if __name__ == "__main__":
    main()
