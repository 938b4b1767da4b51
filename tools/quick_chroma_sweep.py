"""Timing only: python tools/quick_chroma_sweep.py [tracks]  (env AFS_CHROMA_TC_FB / AFS_CHROMA_TC_RING select the split)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

ch = g.submodule("chroma")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = 300 * 22050
argv, sys.argv = sys.argv, sys.argv[:1]
import bench
sys.argv = argv
plan = ch.default_plan()
audio = bench.synth_audio_tracks(torch, T, n, 1000, "cuda").reshape(-1)
offs = np.arange(T + 1, dtype=np.int64) * n
for comp in sys.argv[2:] or ["tc"]:
    out, foffs = plan.run(audio, offs, compute=comp)
    torch.cuda.synchronize()
    frames = int(foffs[-1])
    plan.run(audio, offs, d_out=out, compute=comp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan.run(audio, offs, d_out=out, compute=comp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(comp, "fb", os.environ.get("AFS_CHROMA_TC_FB"), "ring", os.environ.get("AFS_CHROMA_TC_RING"), "frames", frames, "ms", round(ms, 3),
          "Mframes/s", round(frames / ms / 1e3, 2), "HBM frac", round(frames * 8240 / ms / 1e6 / 6545.6, 3), flush=True)
