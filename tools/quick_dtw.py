"""Scratch timing of K2/K3 (not the contract bench): python tools/quick_dtw.py [pairs] [len] [dtype]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
dtw = g.submodule("dtw")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
dt = sys.argv[3] if len(sys.argv) > 3 else "fp64"
plan = dtw.DtwPlan([L] * P, [L] * P, dtype=dt)
tdt = plan.torch_dtype
gen = torch.Generator(device="cuda").manual_seed(1)
a = torch.rand((P, 12, L), device="cuda", dtype=tdt, generator=gen)
b = torch.rand((P, 12, L), device="cuda", dtype=tdt, generator=gen)
a /= a.norm(dim=1, keepdim=True); b /= b.norm(dim=1, keepdim=True)
print("workspace GB", plan.workspace_bytes / 1e9)
for name, fn in (("accumulate", lambda: plan.accumulate(a, b)), ("backtrack", plan.backtrack)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(name, dt, "ms", round(ms, 3), "GCUPS", round(plan.cells / ms / 1e6, 2))
print("path lens", plan.path_len.cpu().numpy()[:4], "acc_end", plan.acc_end.cpu().numpy()[:2])
