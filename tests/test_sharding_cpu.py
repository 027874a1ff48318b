"""Host-side multi-GPU logic on CPU: partition plans, and the time-shard loudness protocol (halo hand-off +
histogram all-reduce) run for real over torch.distributed/gloo with world_size 2, the oracle standing in for
the kernels.  The all-reduced histogram must equal the single-process oracle histogram."""
import os
import socket

import numpy as np
import pytest


def test_shard_tracks_covers_everything():
    from audio_mastering_engine_b200 import sharding
    for n, w in ((1024, 8), (10, 4), (3, 8), (0, 2)):
        got = [i for r in range(w) for i in sharding.shard_tracks(n, w, r)]
        assert got == list(range(n))
        sizes = [len(sharding.shard_tracks(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_time_shard_plan_and_halo():
    from audio_mastering_engine_b200 import sharding, design
    fs = 96000
    n = 3600 * fs
    spans = sharding.plan_time_shards(n, fs, 8)
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all((e - b) == 15 * 30 * fs for b, e in spans)          # C3: 120 chunks -> 15 per GPU
    for f in (44100, 48000, 96000, 192000):
        h = sharding.halo_frames(f)
        s100 = sharding.sub_block_frames(f)
        assert h % s100 == 0 and h % 8 == 0 and h >= design.kw_warm_frames(f) + 3 * s100
    ragged = sharding.plan_time_shards(100 * 48000 + 17, 48000, 3)
    assert ragged[-1][1] == 100 * 48000 + 17 and [e - b for b, e in ragged][:2] == [2 * 30 * 48000, 30 * 48000]
    with pytest.raises(ValueError):
        sharding.plan_time_shards(10 ** 6, 11025, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fs, seconds, chunk_seconds, out_dir):
    import torch
    import torch.distributed as dist
    from audio_mastering_engine_b200 import sharding, synth
    from oracle import chain
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    settings = dict(synth.c2_settings())
    track = synth.track(seconds, fs, track_id=4, am_hz=1.0, drift_db=8.0, drift_period=2.0)
    spans = sharding.plan_time_shards(len(track), fs, world, chunk_seconds)
    b, e = spans[rank]
    pre = np.concatenate([chain.process_chunk(track[s:t], fs, settings)
                          for s, t in [(b + x, min(b + x + int(chunk_seconds * fs), e))
                                       for x in range(0, e - b, int(chunk_seconds * fs))]]) if e > b else np.zeros((0, 2), np.int16)
    halo = sharding.halo_frames(fs)
    tail = np.zeros((halo, 2), np.int16)
    k = min(halo, len(pre))
    if k:
        tail[halo - k:] = pre[len(pre) - k:]
    tails = [torch.zeros((halo, 1), dtype=torch.int32) for _ in range(world)]    # a stereo frame as one int32
    dist.all_gather(tails, torch.from_numpy(tail).view(torch.int32))
    tails = [t.view(torch.int16) for t in tails]
    s100 = sharding.sub_block_frames(fs)
    if rank > 0 and e > b:
        local = np.concatenate([tails[rank - 1].numpy(), pre])
        first_block = halo // s100 - 3
    else:
        local, first_block = pre, 0
    blocks, _ = chain.gating_block_energies(local, fs)
    hist = torch.from_numpy(chain.block_histogram(blocks[first_block:]))
    dist.all_reduce(hist)
    if rank == 0:
        np.save(os.path.join(out_dir, "hist.npy"), hist.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_time_shard_histogram_allreduce_gloo(tmp_path):
    import torch.multiprocessing as mp
    from audio_mastering_engine_b200 import synth
    from oracle import chain
    fs, seconds, chunk_seconds, world = 44100, 6.0, 1.0, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, fs, seconds, chunk_seconds, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "hist.npy")
    track = synth.track(seconds, fs, track_id=4, am_hz=1.0, drift_db=8.0, drift_period=2.0)
    taps = {}
    chain.master(track, fs, synth.c2_settings(), chunk_seconds=chunk_seconds, taps=taps)
    blocks, _ = chain.gating_block_energies(taps["pre_norm"], fs)
    want = chain.block_histogram(blocks)
    assert want.sum() == len(blocks) and want.sum() > 40
    assert np.array_equal(got, want)
