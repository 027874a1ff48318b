"""A/B of the EQ arithmetic: FP64 cascade (the product) against the same cascade in FP32
(ame_plan_options.precision = 1), on the 128-track C4 shard with ONE plan wave (the shape of the ncu capture):
k_eq time per launch, and what the FP32 rounding does to the result - differing samples and the largest difference
at the pre-normalisation stage and at the final output (after the make-up gain of the loudness stage), per loudness
target of the sweep.  Run on a B200: python profiles/scripts/precision_ab.py > gpurun_out/precision_ab.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_mastering_engine_b200 import MasterPlan, synth, EQ_PRESETS  # noqa: E402

dev = torch.device("cuda", 0)
fs, secs, n_tr = 48000, 180.0, int(os.environ.get("AB_TRACKS", "128"))
n = int(secs * fs)
ids = list(range(n_tr))
settings = [synth.c4_settings(t, EQ_PRESETS) for t in ids]
res = {}
d_in = None
for prec in ("exact", "fp32"):
    plan = MasterPlan([n] * n_tr, fs, settings, n_waves=1, precision=prec)
    if d_in is None:
        d_in = torch.zeros((plan.total_frames, 2), dtype=torch.int16, device=dev)
        v = d_in.view(n_tr, -1, 2)
        for k, t in enumerate(ids):
            v[k, :n] = synth.torch_track_batch(1, secs, fs, dev, first_track_id=t)[0]
    d_out = torch.empty_like(d_in)
    for _ in range(2):
        plan.master_device(d_in, d_out, fetch_results=False)
    torch.cuda.synchronize()
    plan.set_timing(True)
    for _ in range(3):
        plan.master_device(d_in, d_out, fetch_results=False)
    torch.cuda.synchronize()
    kt, steps = plan.kernel_times()
    plan.set_timing(False)
    infos = plan.master_device(d_in, d_out)
    pre = plan.read_tap("pre", plan.total_frames * 2, np.int16).reshape(-1, 2).copy()
    res[prec] = dict(k_eq_ms=kt["k_eq"][0] / max(kt["k_eq"][1], 1), step_ms=sum(v[0] for v in kt.values()) / 3,
                     out=d_out.cpu().numpy(), pre=pre, infos=infos)
    plan.close()

a, b = res["exact"], res["fp32"]
print(f"# precision A/B, {n_tr} x {secs:g} s x {fs} Hz C4 tracks, one plan wave (B200)")
print(f"k_eq per launch: FP64 {a['k_eq_ms']:.3f} ms   FP32 {b['k_eq_ms']:.3f} ms   ({a['k_eq_ms'] / b['k_eq_ms']:.2f}x)")
print(f"sum of kernel times per step: FP64 {a['step_ms']:.2f} ms   FP32 {b['step_ms']:.2f} ms")
stride = len(a["out"]) // n_tr
print("\nper loudness target (make-up gain of the sweep's tracks): FP32 result against the FP64 result")
print("lufs  tracks  gain(min..max)  pre: differing samples, max |diff| LSB   out: differing, max |diff| LSB = dBFS   max |LUFS diff|")
for target in (-16.0, -14.0, -9.0):
    ks = [k for k, s in enumerate(settings) if s["lufs"] == target]
    dp = dpn = do = don = 0
    tot = 0
    gl, gh, dl = 1e9, 0.0, 0.0
    for k in ks:
        sl = slice(k * stride, k * stride + n)
        e = np.abs(a["pre"][sl].astype(np.int32) - b["pre"][sl].astype(np.int32))
        f = np.abs(a["out"][sl].astype(np.int32) - b["out"][sl].astype(np.int32))
        dp, dpn = max(dp, int(e.max())), dpn + int((e != 0).sum())
        do, don = max(do, int(f.max())), don + int((f != 0).sum())
        tot += e.size
        g = a["infos"][k]["gain"]
        gl, gh = min(gl, g), max(gh, g)
        dl = max(dl, abs(a["infos"][k]["input_i"] - b["infos"][k]["input_i"]))
    db = 20 * np.log10(max(do, 1e-9) / 32768.0)
    print(f"{target:5.0f}  {len(ks):5d}  {gl:5.2f}..{gh:5.2f}   {dpn / tot:9.2e}  {dp:3d}        {don / tot:9.2e}  {do:3d} = {db:6.1f} dBFS   {dl:.2e}")
mb = [k for k, s in enumerate(settings) if s["multiband"]]
nb = [k for k, s in enumerate(settings) if not s["multiband"]]
for name, ks in (("multiband tracks", mb), ("plain tracks", nb)):
    m = 0
    for k in ks:
        sl = slice(k * stride, k * stride + n)
        m = max(m, int(np.abs(a["out"][sl].astype(np.int32) - b["out"][sl].astype(np.int32)).max()))
    print(f"{name}: max |diff| at the output {m} LSB = {20 * np.log10(max(m, 1e-9) / 32768.0):.1f} dBFS (bar: -80 dBFS = 3.27 LSB)")
