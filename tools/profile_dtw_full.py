import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
P_ARG = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sys.argv = sys.argv[:1]
import bench
dtw = g.submodule("dtw")
P, L = P_ARG, 20000            # default: the bench launch (BASELINE config[2], 256 pairs)
live, ref = bench.synth_chroma_pairs(P, L, 2000)
plan = dtw.DtwPlan([L] * P, [L] * P, dtype="fp64")
a = torch.from_numpy(live).cuda(); b = torch.from_numpy(ref).cuda()
for _ in range(2):
    plan.accumulate(a, b); plan.backtrack()
torch.cuda.synchronize()
print("done")
