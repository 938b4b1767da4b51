"""Scratch: sustained DTW throughput with two batches in flight (two plans, two streams) vs one at a time.
python tools/quick_dtw_stream.py [pairs] [len] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
dtw = g.submodule("dtw")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
gen = torch.Generator(device="cuda").manual_seed(1)
a = torch.rand((P, 12, L), device="cuda", dtype=torch.float64, generator=gen)
b = torch.rand((P, 12, L), device="cuda", dtype=torch.float64, generator=gen)
a /= a.norm(dim=1, keepdim=True); b /= b.norm(dim=1, keepdim=True)
plans = [dtw.DtwPlan([L] * P, [L] * P, dtype="fp64") for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
cells = plans[0].cells

def run(nplans):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    for k in range(K):
        q = k % nplans
        with torch.cuda.stream(streams[q]):
            plans[q].accumulate(a, b)
            plans[q].backtrack()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K

for n in (1, 2):
    run(n)
    ms = run(n)
    print("batches in flight", n, "ms/step", round(ms, 3), "GCUPS", round(cells / ms / 1e6, 1), flush=True)
ref = plans[0].path_len.cpu().numpy()[:4]
print("path lens", ref, plans[1].path_len.cpu().numpy()[:4])
