#!/usr/bin/env python
"""CPU experiment behind DESIGN.md section 6 ("scan vs block-Toeplitz"): the BEST case of a tensor-core formulation of the
EQ cascade - exact products (what a 3 x TF32 / BF16 x 9 operand split buys), FP32 accumulation, the cascade's exact
impulse response - against the FP64 recurrence the reference (scipy) runs, compared where the reference truncates:
the int16 pre-normalisation signal (engine.py:254-257).  No GPU needed.
usage: python profiles/scripts/toeplitz_precision.py > profiles/r02/toeplitz_precision.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_mastering_engine_b200 import EQ_PRESETS, synth   # noqa: E402  (settings and the synthetic tracks only)
from oracle import chain                                     # noqa: E402

FS, SECONDS, TAPS = 48000, 20.0, 8192


def run(block):
    what = "one TF32 MMA k step" if block == 8 else "more generous than any MMA"
    print(f"# products exact, exact sums over {block} taps ({what}), FP32 accumulator across these steps")
    print("preset                differing samples   max |diff| LSB   = dBFS    after a make-up gain of 9.76 (C4, -9 LUFS)")
    for tid, (name, eq) in enumerate(EQ_PRESETS.items()):
        pcm = synth.track(SECONDS, FS, track_id=tid)
        x = chain.to_float(pcm)                                        # float32, x / 2**15 exact
        ref = chain.to_pcm(chain.eq(x.copy(), FS, eq))                 # the reference: FP64 recurrence -> float32 -> int16
        imp = np.zeros(TAPS)
        imp[0] = 1.0
        h = chain.eq_channel(imp, FS, eq)                              # FP64 impulse response of the four stages
        out = np.empty_like(x)
        for c in range(2):
            xc = np.concatenate([np.zeros(TAPS - 1), x[:, c].astype(np.float64)])
            acc = np.zeros(x.shape[0], dtype=np.float32)
            for k0 in range(0, TAPS, block):                           # y[n] = sum_k h[k] x[n - k], k in MMA steps
                part = np.zeros(x.shape[0], dtype=np.float64)
                for k in range(k0, k0 + block):
                    part += h[k] * xc[TAPS - 1 - k: TAPS - 1 - k + x.shape[0]]
                acc = (acc + part.astype(np.float32)).astype(np.float32)
            out[:, c] = acc
        got = chain.to_pcm(out)
        d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        mx = int(d.max())
        db = 20 * np.log10(max(mx, 1e-9) / 32768.0)
        print(f"{name:20s}  {np.mean(d > 0):10.3e}        {mx:6d}        {db:7.1f}    {20 * np.log10(max(mx * 9.76, 1e-9) / 32768.0):7.1f} dBFS")


def main():
    print(f"# {SECONDS:.0f} s of synthetic track per preset at {FS} Hz; FIR = first {TAPS} taps of the cascade's FP64 impulse response "
          f"(tail below 1e-15)")
    run(8)
    run(256)
    print("# bar: -80 dBFS = 3.27 LSB at the OUTPUT, i.e. after the make-up gain and the compressor's integer rms thresholds")


if __name__ == "__main__":
    main()
