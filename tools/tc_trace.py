"""Print the pipeline timeline recorded by AFS_CHROMA_TC_TRACE (csrc/chroma_tc.cu): python tools/tc_trace.py file [block] [first_group] [groups]"""
import sys
import numpy as np

path = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g0 = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ng = int(sys.argv[4]) if len(sys.argv) > 4 else 4
raw = open(path, "rb").read()
n_sm, warps, tlen, blocks = np.frombuffer(raw[:16], np.int32)
t = np.frombuffer(raw[16:], np.int64).reshape(n_sm, warps, tlen // 4, 4).astype(np.float64)
t[t == 0] = np.nan
b = t[blk]
C, E1, E2, M = b[0:8], b[8:16], b[16:24], b[24]
L = b[25] if warps > 25 else None
t0 = np.nanmin(b)
last = np.nanmax(b)
n_tiles = int(np.sum(~np.isnan(M[:, 1])))
print("block %d: %d tiles, %.0f cycles total, %.0f cycles per group" % (blk, n_tiles, last - t0, (last - t0) / max(n_tiles / 2, 1)))


def rng(a):
    return "%7.0f..%-7.0f" % (np.nanmin(a) - t0, np.nanmax(a) - t0)


for g in range(g0, g0 + ng):
    for p in range(2):
        T = 2 * g + p
        if L is not None:
            print("          L tile issued %7.0f | C converted %s a1_free %s" % (L[T, 0] - t0, rng(C[:, T, 1]), rng(C[:, T, 2])))
        print("tile %3d  C a1_full %s | M wait_done %7.0f issued %7.0f | E1 d1_ready %s d1_free %s y_full %s" % (
            T, rng(C[:, T, 0]), M[T, 0] - t0, M[T, 1] - t0, rng(E1[:, T, 0]), rng(E1[:, T, 1]), rng(E1[:, T, 2])))
    print("group %2d  M2 wait_done %7.0f issued %7.0f | E2 d2_ready %s d2_free %s stored %s" % (
        g, M[g, 2] - t0, M[g, 3] - t0, rng(E2[:, g, 0]), rng(E2[:, g, 1]), rng(E2[:, g, 2])))
# steady-state period per group over all blocks
per = []
for k in range(blocks):
    m = t[k, 24, :, 3]
    v = m[~np.isnan(m)]
    if len(v) > 6:
        per.append((v[-2] - v[2]) / (len(v) - 4))
print("M2 issue period per group: median %.0f min %.0f max %.0f cycles over %d blocks" % (np.median(per), np.min(per), np.max(per), len(per)))
