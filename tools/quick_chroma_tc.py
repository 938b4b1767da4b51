"""Scratch check + timing of the tcgen05 chroma path: python tools/quick_chroma_tc.py [tracks] [seconds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g
from oracle import afs_oracle as orc

ch = g.submodule("chroma")
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
aud = np.load(os.path.join(GOLD, "audio_15s.npz"))
for tag in ("ref", "live"):
    x = (aud[tag + "_i16"].astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32)
    want = aud[tag + "_chroma"]
    for comp in ("tc", "fp32"):
        got = ch.wav_samples_to_chroma(x, compute=comp)
        print("golden", tag, comp, got.shape, "max err", float(np.abs(got - want).max()), flush=True)
    raw = ch.wav_samples_to_chroma(x, normalize=False, compute="tc")
    wr = aud[tag + "_raw_chroma"]
    print("golden raw", tag, "rel err", float(np.abs(raw - wr).max() / np.abs(wr).max()), flush=True)
rng = np.random.default_rng(3)
t = np.arange(3 * 22050) / 22050
tracks = [(0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32), (0.1 * rng.standard_normal(20001)).astype(np.float32),
          np.zeros(10000, np.float32), (0.3 * rng.standard_normal(2048)).astype(np.float32),
          (3e4 * rng.standard_normal(30000)).astype(np.float32), (1e-6 * rng.standard_normal(30000)).astype(np.float32),
          (0.3 + 1e-3 * np.sin(2 * np.pi * 440 * np.arange(40000) / 22050)).astype(np.float32)]
got = ch.chroma_batch(tracks, compute="tc")
for k, x in enumerate(tracks):
    want = orc.wav_samples_to_chroma(x)
    print("ragged", k, got[k].shape, want.shape, "max err", float(np.abs(got[k] - want).max()) if want.size else None, flush=True)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
n = int(sec * 22050)
argv, sys.argv = sys.argv, sys.argv[:1]
import bench
sys.argv = argv
plan = ch.default_plan()
audio = bench.synth_audio_tracks(torch, T, n, 1000, "cuda").reshape(-1)
offs = np.arange(T + 1, dtype=np.int64) * n
outs = {}
for comp in ("tc", "fp32"):
    out, foffs = plan.run(audio, offs, compute=comp)
    torch.cuda.synchronize()
    frames = int(foffs[-1])
    for _ in range(2):
        plan.run(audio, offs, d_out=out, compute=comp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan.run(audio, offs, d_out=out, compute=comp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    outs[comp] = out.clone()
    print(comp, "frames", frames, "ms", round(ms, 3), "Mframes/s", round(frames / ms / 1e3, 2), "HBM frac", round(frames * 8240 / ms / 1e6 / 6545.6, 3), flush=True)
d = (outs["tc"] - outs["fp32"]).abs()
print("tc vs fp32 on bench audio: max", float(d.max()), "mean", float(d.mean()))
want = orc.wav_samples_to_chroma(audio[:n].cpu().numpy())
got = outs["tc"][: 12 * int(foffs[1])].view(12, -1).cpu().numpy()
print("tc vs oracle track 0: max", float(np.abs(got - want).max()))
