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


class _OracleShard:
    """sharding.TimeShard with the CPU oracle standing in for the CUDA kernels: the same five methods, so that the
    SHIPPED protocol function sharding.time_sharded_step runs unchanged over gloo on a box without a GPU."""

    def __init__(self, track, span, fs, settings, rank, world, chunk_seconds):
        import torch
        from audio_mastering_engine_b200 import sharding
        self.torch, self.fs, self.settings, self.chunk_seconds = torch, fs, settings, chunk_seconds
        self.begin, self.end = span
        self.n = self.end - self.begin
        self.rank, self.world = rank, world
        self.send = sharding.halo_frames(fs)
        self.halo = self.send if (rank > 0 and self.begin > 0 and self.n > 0) else 0
        self.dev = torch.device("cpu")
        self.x = track[self.begin:self.end]
        self.pre = np.zeros((self.halo + self.n, 2), np.int16)

    def pre_normalisation(self):
        from oracle import chain
        cf = int(self.chunk_seconds * self.fs)
        parts = [chain.process_chunk(self.x[s:s + cf], self.fs, self.settings) for s in range(0, self.n, cf)]
        if parts:
            self.pre[self.halo:] = np.concatenate(parts)

    def tail(self):
        out = np.zeros((self.send, 2), np.int16)
        k = min(self.send, self.n + self.halo)
        if k:
            out[self.send - k:] = self.pre[self.halo + self.n - k:self.halo + self.n]
        return self.torch.from_numpy(out)

    def set_halo(self, prev_tail):
        if self.halo:
            self.pre[:self.halo] = prev_tail.numpy()[-self.halo:]

    def histogram(self):
        from audio_mastering_engine_b200 import sharding
        from oracle import chain
        hist = np.zeros((1, 1000), np.int64)
        if self.n:
            blocks, _ = chain.gating_block_energies(self.pre, self.fs)
            first = self.halo // sharding.sub_block_frames(self.fs) - 3 if self.halo else 0
            hist[0] = chain.block_histogram(blocks[first:])
        return self.torch.from_numpy(hist)

    def normalise(self, hist):
        from oracle import chain
        measured, _ = chain.gated_loudness_from_histogram(hist.numpy()[0])
        if not self.n:
            return self.torch.zeros((0, 2), dtype=self.torch.int16), None
        gain, mi = chain.static_gain_from_measured(measured, self.settings["lufs"])
        out = chain.apply_static_gain(self.pre[self.halo:], gain)
        return self.torch.from_numpy(out), dict(input_i=measured, n_blocks=int(hist.sum()))


def _worker(rank, world, port, fs, seconds, chunk_seconds, out_dir):
    import torch.distributed as dist
    from audio_mastering_engine_b200 import sharding, synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    settings = dict(synth.c2_settings())
    track = synth.track(seconds, fs, track_id=4, am_hz=1.0, drift_db=8.0, drift_period=2.0)
    spans = sharding.plan_time_shards(len(track), fs, world, chunk_seconds)
    sh = _OracleShard(track, spans[rank], fs, settings, rank, world, chunk_seconds)
    out, info, moved = sharding.time_sharded_step(sh, spans)       # the shipped protocol: halo send/recv + all-reduce
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out.numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "info.npy"), np.array([info["input_i"], info["n_blocks"], moved]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_time_shard_protocol_gloo(tmp_path, world):
    """sharding.time_sharded_step over real gloo, the oracle standing in for the kernels: the gathered spans must be
    the single-process oracle result bit for bit (global loudness from the all-reduced histogram)."""
    import torch.multiprocessing as mp
    from audio_mastering_engine_b200 import synth
    from oracle import chain
    fs, seconds, chunk_seconds = 44100, 6.0, 1.0
    port = _free_port()
    mp.spawn(_worker, args=(world, port, fs, seconds, chunk_seconds, str(tmp_path)), nprocs=world, join=True)
    track = synth.track(seconds, fs, track_id=4, am_hz=1.0, drift_db=8.0, drift_period=2.0)
    want, winfo = chain.master(track, fs, synth.c2_settings(), chunk_seconds=chunk_seconds)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)], axis=0)
    info = np.load(tmp_path / "info.npy")
    assert info[0] == winfo["input_i"] and int(info[1]) == winfo["n_blocks"] and winfo["n_blocks"] > 40
    assert info[2] > 0
    assert np.array_equal(got, want)
