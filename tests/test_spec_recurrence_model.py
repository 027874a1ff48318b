"""Host-side model of the algorithm inside k_att_chain_spec (csrc/ame_kernels.cuh): the compressor's attenuation
recurrence evaluated in S time segments from guessed starts and repaired until exact.  The model restates the
kernel's control flow in Python - pass 1 from a guessed start, repair passes that carry the old and the new start and stop
where they are equal, starts handed across silent segments - and checks the claim the kernel relies on: when no
segment's start changes any more, every stored value is the one the sequential loop (pydub's, oracle/chain.py)
produces, bit for bit; and the tau form of the update used on the GPU is the reference's update.
CPU only; the CUDA kernel itself is compared with the sequential kernel and the oracle in test_gpu_parity.py."""
import math

import numpy as np
import pytest

from oracle import chain


def _entries(fs, threshold, ratio):
    """(M, inc, dec, tau) per integer rms, as build_att_table in csrc/ame.cu does; flagged(r) = r > thresh_rms."""
    thr, _look, attack_frames, release_frames = chain.compressor_constants(fs, threshold, ratio)
    table = {}

    def entry(r):
        if r not in table:
            m = chain.max_attenuation_for_rms(r, thr, ratio)
            inc, dec = m / attack_frames, m / release_frames
            tau = max(m - inc, 0.0)
            while tau > 0 and tau + inc >= m:
                tau = math.nextafter(tau, -1.0)
            while tau + inc < m:
                tau = math.nextafter(tau, 1e300)
            table[r] = (m, inc, dec, tau)
        return table[r]
    return thr, entry


def _update_ref(att, m, inc, dec):
    """pydub's step for a frame above threshold (below it M = 0 and nothing moves)."""
    if att <= m:
        return min(att + inc, m)
    return max(att - dec, 0)


def _update_gpu(att, m, inc, dec, tau):
    """att_update() of the kernels: both predicates on the OLD attenuation."""
    if att > m:
        return att - dec
    return att + inc if att < tau else m


def _sequential(rms, thr, entry):
    att, out = 0.0, np.zeros(len(rms))
    for i, r in enumerate(rms):
        if r > thr:
            m, inc, dec, _ = entry(int(r))
            att = _update_ref(att, m, inc, dec)
        out[i] = att
    return out


def _walk(rms, thr, entry, b0, b1, b, out, a=None):
    """Frames [b0, b1) from attenuation b, storing into out; with `a` also the previous start's trajectory, stopping
    after the first flagged frame where the two are equal.  Returns (end value, met)."""
    for i in range(b0, b1):
        r = rms[i]
        if r > thr:
            e = entry(int(r))
            b = _update_gpu(b, *e)
            if a is not None:
                a = _update_gpu(a, *e)
        out[i] = b
        if a is not None and r > thr and a == b:
            return b, True
    return b, False


def _speculative(rms, thr, entry, n_seg):
    n = len(rms)
    seg = -(-n // n_seg)
    bounds = [(min(n, t * seg), min(n, (t + 1) * seg)) for t in range(n_seg)]
    out = np.zeros(n)
    start = [0.0] * n_seg
    end = [0.0] * n_seg
    flagged = [bool(np.any(rms[b0:b1] > thr)) for b0, b1 in bounds]
    for t, (b0, b1) in enumerate(bounds):                       # pass 1, from a guess: M of the last flagged frame in front
        if t > 0 and b0 < b1:
            for i in range(b0 - 1, max(0, b0 - 64) - 1, -1):
                if rms[i] > thr:
                    start[t] = entry(int(rms[i]))[0]
                    break
        end[t], _ = _walk(rms, thr, entry, b0, b1, start[t], out)
    prev = []
    for t in range(n_seg):                                      # last segment before t that holds a flagged frame
        u = t - 1
        while u >= 0 and not flagged[u]:
            u -= 1
        prev.append(u)
    passes = 0
    while True:
        frm = [end[prev[t]] if prev[t] >= 0 else 0.0 for t in range(n_seg)]     # all lanes read, then all write
        redo = [t for t in range(n_seg) if frm[t] != start[t]]
        if not redo:
            return out, passes
        passes += 1
        assert passes <= n_seg, "one more segment must become final in every pass"
        for t in redo:
            b0, b1 = bounds[t]
            b, met = _walk(rms, thr, entry, b0, b1, frm[t], out, a=start[t])
            if not met:
                end[t] = b
            start[t] = frm[t]


def _rms_series(rng, n, kind):
    t = np.arange(n)
    if kind == "bursts":                 # loud / quiet alternation: the attenuation clamps and parks in turn
        level = 2500 + 2200 * np.sign(np.sin(2 * np.pi * t / 1700.0)) + rng.integers(-300, 300, n)
    elif kind == "held":                 # one loud passage, then a bed just over threshold: never clamps again
        level = np.where(t < 400, 20000, 3320) + rng.integers(-15, 15, n)
    elif kind == "sparse":               # a handful of flagged frames in mostly silent segments
        level = np.where(rng.random(n) < 0.004, 9000, 100)
    else:                                # noise around the threshold
        level = 3277 + rng.integers(-1500, 1500, n)
    return np.clip(level, 0, 32768).astype(np.int64)


@pytest.mark.parametrize("kind", ["bursts", "held", "sparse", "noise"])
@pytest.mark.parametrize("n_seg", [1, 7, 32])
def test_speculate_and_repair_equals_sequential(kind, n_seg):
    rng = np.random.default_rng(11)
    fs, threshold, ratio = 48000, -20.0, 4.0
    thr, entry = _entries(fs, threshold, ratio)
    rms = _rms_series(rng, 6000, kind)
    want = _sequential(rms, thr, entry)
    got, passes = _speculative(rms, thr, entry, n_seg)
    assert np.array_equal(got, want), (kind, n_seg)
    if kind == "held" and n_seg > 1:
        assert passes >= n_seg // 2          # the strictly sequential case: about one segment settles per pass


def test_tau_form_is_the_reference_update():
    """att < tau  <=>  fl(att + inc) < M for the tau of the table, so both forms take the same branch everywhere."""
    rng = np.random.default_rng(5)
    thr, entry = _entries(44100, -25.0, 6.0)
    for r in rng.integers(int(thr) + 1, 32768, 300):
        m, inc, dec, tau = entry(int(r))
        probes = [0.0, tau, math.nextafter(tau, 0.0), math.nextafter(tau, 1e9), m, math.nextafter(m, 1e9), m - inc, 2 * m]
        probes += list(rng.uniform(0, 1.5 * m, 20))
        for att in probes:
            assert _update_gpu(att, m, inc, dec, tau) == _update_ref(att, m, inc, dec), (r, att)
