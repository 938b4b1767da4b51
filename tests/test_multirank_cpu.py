"""CPU suite, world_size = 2 over gloo: host-side logic of the multi-GPU paths — unit sharding
used by bench.py / batch front-ends and the right-to-left backtrack hand-off of the striped DTW."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, full_path, bounds, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    striped = g.submodule("striped")
    c0, c1 = bounds[rank]

    def local_backtrack(start_row):
        # stand-in for the stripe kernel: the part of a known global path inside this stripe
        seg = full_path[(full_path[:, 1] >= c0) & (full_path[:, 1] < c1)]
        assert seg[-1, 0] == start_row, (rank, seg[-1], start_row)
        if rank == 0:
            return seg, -1
        first = seg[0]
        k = int(np.nonzero((full_path == first).all(axis=1))[0][0])
        return seg, int(full_path[k - 1, 0])

    group = dist.new_group(list(range(world)))           # an explicit group: ranks are passed through it everywhere
    seg = striped.handoff_backtrack(rank, world, local_backtrack, int(full_path[-1, 0]), dist, group=group)
    gathered = striped.gather_segments(seg, rank, world, dist, group=group)
    if rank == 0:
        np.save(os.path.join(out_dir, "path.npy"), striped.stitch_segments(gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_striped_backtrack_handoff_world2(tmp_path, orc):
    rng = np.random.default_rng(0)
    a, b = rng.random((12, 90)), rng.random((12, 131))
    _, _, path = orc.DTW(a, b, dense=False)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    striped = g.submodule("striped")
    bounds = striped.stripe_bounds(131, 2)
    assert bounds == [(0, 66), (66, 131)]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, path, bounds, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "path.npy"))
    assert np.array_equal(got, path)


def test_stripe_bounds_and_stitch(entry):
    striped = entry.submodule("striped")
    assert striped.stripe_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert striped.stripe_bounds(2, 8) == [(0, 1), (1, 2)]
    b = striped.stripe_bounds(200000, 8)
    assert b[0] == (0, 25000) and b[-1] == (175000, 200000) and all(c1 - c0 == 25000 for c0, c1 in b)
    segs = [np.array([[0, 0], [1, 1]]), np.empty((0, 2), dtype=np.int64), np.array([[2, 5]])]
    assert striped.stitch_segments(segs).tolist() == [[0, 0], [1, 1], [2, 5]]
