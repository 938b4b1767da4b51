"""Multi-GPU check + timing of the striped DTW (K4): torchrun --nproc-per-node G tools/striped_dist.py [M] [N] [check]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as g

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
striped = g.submodule("striped")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
check = len(sys.argv) > 3 and sys.argv[3] == "check"
sys.argv = sys.argv[:1]
import bench
live, ref = bench.synth_chroma_pairs(1, max(M, N), 5000)
a, b = np.ascontiguousarray(live[0][:, :M]), np.ascontiguousarray(ref[0][:, :N])
sd = striped.StripedDtwDistributed(M, N, dist)
c0, c1 = sd.bounds[rank]
d_a = torch.from_numpy(a).cuda()
d_b = torch.from_numpy(np.ascontiguousarray(b[:, c0:c1])).cuda()
times = []
for it in range(3):
    sd.reset()
    t0 = time.perf_counter()
    sd.accumulate(d_a, d_b)
    torch.cuda.synchronize()
    dist.barrier()
    times.append(time.perf_counter() - t0)
t0 = time.perf_counter()
path = sd.backtrack()
t_bt = time.perf_counter() - t0
end = sd.acc_end()
end_all = torch.tensor([end if end is not None else 0.0], dtype=torch.float64, device="cuda")
dist.all_reduce(end_all)            # only the last rank contributes a non-zero value
if rank == world - 1:
    print("rank", rank, "acc_end", end, "accumulate s", [round(t, 4) for t in times], "GCUPS", round(M * N / min(times) / 1e9, 1), flush=True)
if rank == 0:
    print("path", path.shape, path[0], path[-1], "backtrack s", round(t_bt, 4), flush=True)
    if check:
        from oracle import afs_oracle as orc
        _, oend, opath = orc.DTW(a, b, dense=False)
        print("PARITY path", bool(np.array_equal(path, opath)), "acc_end", bool(float(end_all.item()) == oend), "oracle acc_end", oend, flush=True)
sd.close()
dist.destroy_process_group()
