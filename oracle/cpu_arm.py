"""Timed CPU arm for bench.py (cpu_baseline / --impl reference).  TEST INFRASTRUCTURE (oracle/README.md).

Runs the oracle port the way the reference runs its own path: per-chunk numpy/scipy stages plus the
per-frame Python loop of pydub's compressor on real audioop (NOT the C twin - the C twin exists only to
make the checker fast; the reference's cost is the Python loop), then the loudness restatement.
The ffmpeg stages (split, concat, loudnorm, alimiter) cannot be timed - the binary is absent - and are
substituted by the numpy restatements, which is stated in the bench output.
"""
from __future__ import annotations

import os
import time

import numpy as np


def _one(args):
    """Master one synthetic sample on one core.  Returns (audio_seconds, wall_seconds)."""
    track_id, seconds, fs = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from audio_mastering_engine_b200 import synth
    from oracle import chain
    x = synth.track(seconds, fs, track_id=track_id, am_hz=2.0)
    settings = synth.c4_settings(track_id, chain.EQ_PRESETS)
    t0 = time.perf_counter()
    chain.master(x, fs, settings, compress=chain.compress_dynamic_range_py)
    return seconds, time.perf_counter() - t0


def run_step(pool, track_ids, seconds, fs):
    """One bounded sample of the C4 workload: len(track_ids) tracks of `seconds` audio, one per worker.
    Returns (audio seconds mastered, wall seconds)."""
    t0 = time.perf_counter()
    if pool is None:
        res = [_one((t, seconds, fs)) for t in track_ids]
    else:
        res = list(pool.map(_one, [(t, seconds, fs) for t in track_ids]))
    wall = time.perf_counter() - t0
    return float(sum(r[0] for r in res)), wall


def make_pool(workers):
    if workers <= 1:
        return None
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    return ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn"))
