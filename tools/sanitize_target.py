"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
rng = np.random.default_rng(0)
dtw, batch, chroma, striped = (g.submodule(n) for n in ("dtw", "batch", "chroma", "striped"))
a, b = rng.random((12, 300)), rng.random((12, 333))
dtw.DTW(a, b)                                   # dense single pair
dtw.DTW(a, b, dtype="fp32")
dtw.dtw_batch([rng.random((12, m)) for m in (129, 40, 260)], [rng.random((12, n)) for n in (65, 200, 31)])
striped.dtw_striped_local(a, b, 3)
for kind in ("otw", "livenote_v2"):
    ob = batch.OtwBatch([b, a[:, :200]], 24, 3, kind=kind)
    fr = torch.from_numpy(np.ascontiguousarray(np.stack([a.T[:150], a.T[:150]], axis=1))).cuda()
    ob.step_device(fr)
    ob.paths(); ob.close()
wb = batch.WtwBatch([b, b], 12, 5)
wb.push(np.ascontiguousarray(np.stack([a.T[:120], a.T[:120]], axis=1)))
wb.paths(); wb.close()
x = (0.3 * rng.standard_normal(30000)).astype(np.float32)
chroma.chroma_batch([x, x[:9001], x[:100]])
chroma.chroma_batch([x[:20000]], compute="fp64")
torch.cuda.synchronize()
print("sanitize target done")
