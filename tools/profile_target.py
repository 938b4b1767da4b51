"""Small fixed workload for ncu captures: python tools/profile_target.py [dtw|chroma|otw|all]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
what = sys.argv[1] if len(sys.argv) > 1 else "all"
sys.argv = sys.argv[:1]
import bench
if what in ("dtw", "all"):
    dtw = g.submodule("dtw")
    P, L = 32, 8000
    live, ref = bench.synth_chroma_pairs(P, L, 2000)
    plan = dtw.DtwPlan([L] * P, [L] * P, dtype="fp64")
    a = torch.from_numpy(live).cuda(); b = torch.from_numpy(ref).cuda()
    for _ in range(3):
        plan.accumulate(a, b); plan.backtrack()
    torch.cuda.synchronize()
if what in ("chroma", "all"):
    ch = g.submodule("chroma")
    T, n = 64, 60 * 22050
    audio = bench.synth_audio_tracks(torch, T, n, 1000, "cuda").reshape(-1)
    offs = np.arange(T + 1, dtype=np.int64) * n
    plan = ch.default_plan()
    out, _ = plan.run(audio, offs)
    for _ in range(2):
        plan.run(audio, offs, d_out=out)
    torch.cuda.synchronize()
if what in ("otw", "all"):
    batch = g.submodule("batch")
    ref, frames = bench.synth_streams(torch, 4096, 3000, 760, 3000, "cuda")
    b = batch.OtwBatch(ref, 500, 3, kind="otw")
    b.step_device(frames[:700].contiguous(), want_points=False)     # one launch: reach steady state (t >= c)
    for k in range(700, 705):
        b.step_device(frames[k])
    torch.cuda.synchronize()
print("done", what)
