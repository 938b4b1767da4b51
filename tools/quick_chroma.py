"""Scratch timing of K1: python tools/quick_chroma.py [tracks] [seconds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
ch = g.submodule("chroma")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
n = int(sec * 22050)
plan = ch.default_plan()
audio = torch.randn(T * n, device="cuda", dtype=torch.float32) * 0.1
offs = np.arange(T + 1, dtype=np.int64) * n
out, foffs = plan.run(audio, offs)
torch.cuda.synchronize()
frames = int(foffs[-1])
for comp in ("fp32", "fp64"):
    for _ in range(2): plan.run(audio, offs, d_out=out, compute=comp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): plan.run(audio, offs, d_out=out, compute=comp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(comp, "frames", frames, "ms", round(ms, 3), "Mframes/s", round(frames / ms / 1e3, 2), "GB/s(alg)", round(frames * 8240 / ms / 1e6, 1))
