#!/usr/bin/env python
"""Benchmark of the alignment hot path (contract: see the task statement / DESIGN.md §6).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port on host cores)

Headline metric (BASELINE.json): DTW GCUPS on config[2] — 20k x 20k chroma frames,
256 pairs sharded over 8 GPUs = 32 pairs per GPU (weak scaling: 32 pairs per rank).
One "step" = one pass (accumulate K2 + backtrack K3) over the rank's 32 pairs.
The same JSON line carries the other two metrics BASELINE.json names as sub-objects
(`chroma`: frames/s on config[1]; `otw`: p99 per-frame latency on config[3]) when
those workloads are enabled (--workloads dtw,chroma,otw; default all available).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DTW_LEN = 20000
DTW_PAIRS_PER_GPU = 32          # 256 pairs / 8 GPUs (BASELINE.json configs[2])
FP64_LANES_PER_SM = 64          # B200 FP64 FMA lanes per SM
DTW_FLOP_PER_CELL = 31          # 12 FMA + sub + scale + 3 add + 2 cmp/sel


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------------------- synthetic data
def synth_chroma_pairs(n_pairs, length, seed0):
    """SURVEY.md §8(d) cfg3 generator: AR(1)-smoothed random chroma, unit columns;
    live = ref sampled along a smooth monotone warp + noise.  Returns (live, ref)
    as float64 arrays (n_pairs, 12, length)."""
    ref = np.empty((n_pairs, 12, length))
    live = np.empty((n_pairs, 12, length))
    u = np.linspace(0.0, 1.0, length)
    warp = np.clip((u + 0.08 * np.sin(6 * np.pi * u)), 0, 1) * (length - 1)
    idx = np.round(warp).astype(np.int64)
    for p in range(n_pairs):
        rng = np.random.default_rng(seed0 + p)
        x = rng.random((12, length))
        # x[:,k] = 0.7 x[:,k-1] + 0.3 x[:,k]  (vectorised as a truncated exponential filter)
        taps = 0.3 * 0.7 ** np.arange(64)
        y = np.empty_like(x)
        for f in range(12):
            y[f] = np.convolve(x[f], taps)[:length]
        y[:, 0] = x[:, 0]
        y /= np.linalg.norm(y, axis=0)
        ref[p] = y
        z = y[:, idx] + 0.05 * rng.random((12, length))
        z /= np.linalg.norm(z, axis=0)
        live[p] = z
    return live, ref


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = float(s[1])
            except ValueError:
                continue
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # the busiest samples are the ones under load: take the median of the upper half
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_dtw_baseline(seconds=12.0, threads=None):
    """Oracle C port (oracle/afs_oracle.c, dtw.py:5-53 arithmetic) on the host cores:
    one 3000 x 3000 pair per call, calls spread over `threads` Python threads
    (ctypes releases the GIL) until `seconds` elapse."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from oracle import afs_oracle as orc
    L = orc.lib()
    threads = threads or os.cpu_count() or 1
    n = 3000
    live, ref = synth_chroma_pairs(1, n, 2000)
    a = np.ascontiguousarray(live[0])
    b = np.ascontiguousarray(ref[0])
    t_end = time.perf_counter() + seconds
    counts = [0] * threads

    def work(k):
        while time.perf_counter() < t_end:
            L.orc_dtw_many(a.ctypes.data_as(orc._f64p), b.ctypes.data_as(orc._f64p), 12, n, n, 1, None)
            counts[k] += 1

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    cells = sum(counts) * n * n
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
            "sample": "%d x (3000x3000 pair, fp64, oracle/afs_oracle.c) over %d threads in %.1f s" % (sum(counts), threads, dt)}


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    vals = []
    base = None
    for it in range(args.warmup + args.steps):
        base = cpu_dtw_baseline(seconds=per_step)
        if it >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {
        "impl": "reference", "metric": "dtw_gcups", "value": v, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "offline full DTW 20k x 20k chroma frames, 32 pairs per GPU (cfg[2]); CPU arm = bounded sample of 3000x3000 pairs, same arithmetic"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workloads", default="all")
    ap.add_argument("--dtype", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--pairs", type=int, default=DTW_PAIRS_PER_GPU)
    ap.add_argument("--length", type=int, default=DTW_LEN)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup_effective = 3
    else:
        args.warmup_effective = args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import __graft_entry__ as g

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nat = g.submodule("_native")
    dtw = g.submodule("dtw")
    wl = ["dtw", "chroma", "otw"] if args.workloads == "all" else args.workloads.split(",")

    # ---------------- DTW (headline) ----------------
    P, Ln = args.pairs, args.length
    live, ref = synth_chroma_pairs(P, Ln, 2000 + rank * P)
    npdt = np.float64 if args.dtype == "fp64" else np.float32
    h_a = torch.from_numpy(np.ascontiguousarray(live, dtype=npdt)).pin_memory()
    h_b = torch.from_numpy(np.ascontiguousarray(ref, dtype=npdt)).pin_memory()
    plan = dtw.DtwPlan([Ln] * P, [Ln] * P, dtype=args.dtype)
    d_a = h_a.to("cuda")
    d_b = h_b.to("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step_resident():
        plan.accumulate(d_a, d_b)
        plan.backtrack()

    for _ in range(args.warmup_effective):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = nat.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the events)
        ev[k][0].record()
        plan.accumulate(d_a, d_b)
        ev[k][1].record()
        plan.backtrack()
        ev[k][2].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = nat.launch_count() - launches0
    acc_ms = [e[0].elapsed_time(e[1]) for e in ev]
    bt_ms = [e[1].elapsed_time(e[2]) for e in ev]
    step_ms = float(np.mean(acc_ms) + np.mean(bt_ms))
    step_ms_max = max_over_ranks(step_ms)
    cells_rank = float(plan.cells)
    value = cells_rank * world / (step_ms_max * 1e-3) / 1e9

    # ---------------- e2e: host buffers through the public API objects ----------------
    e2e_steps = max(2, min(args.steps, 5))
    h_start = torch.empty(P, dtype=torch.int32).pin_memory()
    h_len = torch.empty(P, dtype=torch.int32).pin_memory()
    h_path = torch.empty((plan.path_total, 2), dtype=torch.int32).pin_memory()
    h_end = torch.empty(P, dtype=torch.float64).pin_memory()

    def step_e2e():
        da = h_a.to("cuda", non_blocking=True)
        db = h_b.to("cuda", non_blocking=True)
        plan.accumulate(da, db)
        plan.backtrack()
        h_start.copy_(plan.path_start, non_blocking=True)
        h_len.copy_(plan.path_len, non_blocking=True)
        h_path.copy_(plan.path, non_blocking=True)
        h_end.copy_(plan.acc_end, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_val = cells_rank * world / e2e_s / 1e9
    h2d = int(h_a.numel() * h_a.element_size() + h_b.numel() * h_b.element_size())
    d2h = int(h_path.numel() * 4 + h_start.numel() * 4 + h_len.numel() * 4 + h_end.numel() * 8)

    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    lanes = FP64_LANES_PER_SM if args.dtype == "fp64" else 128
    pipe_peak_tflops = n_sm * lanes * 2 * sm_max * 1e6 / 1e12
    acc_ms_mean = float(np.mean(acc_ms))
    achieved_tflops = cells_rank * DTW_FLOP_PER_CELL / (acc_ms_mean * 1e-3) / 1e12
    alg_bytes = cells_rank * 0.25 + 2 * P * 12 * Ln * (8 if args.dtype == "fp64" else 4)
    roofline = {
        "kernel": "dtw_wavefront_kernel<%s>" % ("double" if args.dtype == "fp64" else "float"),
        "bound": "fp64_pipe" if args.dtype == "fp64" else "fp32_pipe",
        "achieved": achieved_tflops, "peak": pipe_peak_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / pipe_peak_tflops,
        "peak_source": "derived: %d SMs x %d lanes x 2 x %.0f MHz (no %s pipe figure in MEASURED_PEAKS.json)" % (n_sm, lanes, sm_max, args.dtype),
        "flop_per_cell": DTW_FLOP_PER_CELL, "kernel_ms": acc_ms_mean,
        "hbm": {"achieved": alg_bytes / (acc_ms_mean * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                "frac": alg_bytes / (acc_ms_mean * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
                "peak_source": "measured" if "hbm_gbs" in peaks else "fallback"},
        "traffic": None,
    }
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_dtw_baseline(seconds=args.cpu_seconds)
    line = {
        "metric": "dtw_gcups", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup_effective, "ms_per_step": step_ms_max, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if args.dtype == "fp64" else "f32", "data": "synthetic",
        "config": {"workload": "offline full DTW %dx%d chroma frames, %d pairs per GPU (BASELINE cfg[2]: 256 pairs over 8 GPUs)" % (Ln, Ln, P),
                   "pairs_per_gpu": P, "frames": Ln, "features": 12,
                   "l2": "256 MiB flush between timed steps; per-step working set 3.2 GB direction map > 126 MB L2",
                   "step": "accumulate (K2) + backtrack (K3)", "parallelism": "pairs sharded, no collective"},
        "kernel_ms": {"accumulate": acc_ms_mean, "backtrack": float(np.mean(bt_ms))},
        "wall_s_timed_region": t_wall,
        "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
