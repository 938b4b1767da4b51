#!/usr/bin/env python
"""Benchmark of the alignment hot path (contract: see the task statement / DESIGN.md §6).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port on host cores)

Headline metric (BASELINE.json): DTW GCUPS on config[2] — 20k x 20k chroma frames,
batch of 256 pairs.  The batch fits one GPU (26.5 GB of direction maps), so every GPU runs the configuration in full:
256 pairs per rank at every N (weak scaling).  One "step" = one pass (accumulate K2 + backtrack K3) over the rank's 256
pairs.  (`--pairs 32` gives the config's 8-GPU share per GPU: 492 instead of 640 GCUPS per GPU, see DESIGN.md.)
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
DTW_PAIRS_PER_GPU = 256         # BASELINE.json configs[2]: a batch of 256 pairs; it fits one GPU, so each GPU runs it in full
FP64_LANES_PER_SM = 64          # B200 FP64 FMA lanes per SM
DTW_FLOP_PER_CELL = 31          # 12 FMA + sub + scale + 3 add + 2 cmp/sel


def bind_to_gpu_numa(torch, local, world):
    """Pin this rank to cores of the NUMA node its GPU hangs off, before any pinned host buffer is allocated (first touch
    then places the buffers on that node): with every rank on node 0, the ranks of the far GPUs stage their uploads through
    the inter-socket link.  Returns what was done, for the JSON line."""
    info = {"numa_node": None, "cpus": None}
    try:
        props = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = []
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return info
        # ranks whose GPUs share the node take disjoint slices of its cores
        per_node = max(1, world // max(1, len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])))
        k = local % per_node
        share = allowed[k * len(allowed) // per_node:(k + 1) * len(allowed) // per_node] or allowed
        os.sched_setaffinity(0, share)
        info = {"numa_node": node, "cpus": "%d-%d (%d cores)" % (share[0], share[-1], len(share))}
    except Exception as exc:
        info["error"] = repr(exc)[:120]
    return info


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
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # same workload as the GPU arm (its `config`, `metric`, `unit`); a step here is a bounded SAMPLE of it — the
        # 256 x 20k x 20k batch would take the host cores the better part of an hour per step
        "config": {"workload": "offline full DTW %dx%d chroma frames, one batch of %d pairs sharded over %d GPU(s): %d pairs per GPU (BASELINE cfg[2]; strong scaling)"
                               % (args.length, args.length, args.pairs, max(1, args.gpus), max(1, args.pairs // max(1, args.gpus))),
                   "pairs_total": args.pairs, "frames": args.length, "features": 12,
                   "cpu_sample": "each step times 3000x3000 pairs of the same generator for a fixed wall time on all host threads (GCUPS does not depend on the pair size on the CPU: no wavefront ramp)"},
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
    ap.add_argument("--chroma-tracks", type=int, default=CHROMA_TRACKS)
    ap.add_argument("--chroma-compute", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--otw-streams", type=int, default=OTW_STREAMS)
    ap.add_argument("--otw-steps", type=int, default=0)
    ap.add_argument("--wtw-streams", type=int, default=1024)
    ap.add_argument("--striped-cols-per-gpu", type=int, default=25000)
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
    affinity = bind_to_gpu_numa(torch, local, world) if world > 1 else None
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
    ctx = {"args": args, "rank": rank, "world": world, "local": local, "barrier": barrier, "max_over_ranks": max_over_ranks,
           "nat": nat, "g": g, "torch": torch}
    wl = ["dtw", "chroma", "otw"] if args.workloads == "all" else args.workloads.split(",")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    results = {}
    if "dtw" in wl:
        results["dtw"] = bench_dtw(ctx)
        torch.cuda.empty_cache()          # 80+ GB of plans go back to the driver before the next workload allocates
    if "chroma" in wl:
        results["chroma"] = bench_chroma(ctx)
    if "otw" in wl:
        results["otw"] = bench_otw(ctx)
    if "wtw" in wl or args.workloads == "all":
        results["wtw"] = bench_wtw(ctx)
    if world > 1 and ("striped" in wl or args.workloads == "all"):
        ctx["dist"] = dist
        try:
            results["striped"] = bench_striped(ctx)
        except Exception as exc:         # supplementary line: never lose the headline numbers over it
            results["striped"] = {"metric": "striped_dtw_gcups", "error": repr(exc)[:300]}
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        head_key = "dtw" if "dtw" in results else list(results)[0]
        line = dict(results[head_key])
        for k, v in results.items():
            if k != head_key:
                line[k] = v
        line["clocks"] = clocks
        line["host_affinity_rank0"] = affinity
        # the figures of the other workloads once more as flat scalars, so that they survive any post-processing that keeps
        # only the top level of the line
        def pick(key, *path):
            o = results.get(key)
            for k in path:
                o = o.get(k) if isinstance(o, dict) else None
            return o
        line["chroma_frames_per_s"] = pick("chroma", "value")
        line["chroma_hbm_frac"] = pick("chroma", "roofline", "frac")
        line["chroma_e2e_frames_per_s"] = pick("chroma", "e2e", "value")
        line["chroma_e2e_pcm16_frames_per_s"] = pick("chroma", "e2e_pcm16", "value")
        line["otw_p99_ms"] = pick("otw", "value")
        line["otw_p50_ms"] = pick("otw", "ms_per_step")
        line["wtw_value"] = pick("wtw", "value")
        line["striped_gcups"] = pick("striped", "value")
        line["striped_ms"] = pick("striped", "ms_per_step")
        line["striped_single_gpu_ms"] = pick("striped", "single_gpu_ms")
        line["striped_parity"] = pick("striped", "parity_vs_single_gpu")
        line["dtw_weak_gcups"] = pick("dtw", "weak", "value")
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- WTW (cfg[4]'s windowed variant)
def bench_wtw(ctx):
    """Windowed time warping on chroma columns, many replicas per GPU (WTW is a serial chain of windows per stream, so
    it shards by stream only).  Two window shapes: the offline parameters of test_simple.py:168 (W = 20, h = 10) and
    the live app's (wtw_live.py:106: W = 100, h = 50).  Whole live sequence in one launch; frames/s = live frames of
    all streams / device time."""
    args, rank, world, torch, g = ctx["args"], ctx["rank"], ctx["world"], ctx["torch"], ctx["g"]
    batch = g.submodule("batch")
    S = args.wtw_streams
    ref, frames = synth_streams(torch, S, OTW_REF, OTW_LIVE, 4000 + rank, "cuda")
    out = {"metric": "wtw_frames_per_s", "unit": "live frames/s", "n_gpus": world, "streams_per_gpu": S, "ref_frames": OTW_REF, "shapes": {}}
    for W, h in ((20, 10), (100, 50)):
        b = batch.WtwBatch(ref, W, h)
        b.push_device(frames[:200].contiguous())
        b.reset()
        torch.cuda.synchronize()
        ctx["barrier"]()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = b.push_device(frames)
        e1.record()
        torch.cuda.synchronize()
        ms = ctx["max_over_ranks"](e0.elapsed_time(e1))
        n_live = int(frames.shape[0])
        lens = [len(p) for p in b.paths()[:8]]
        out["shapes"]["W%d_h%d" % (W, h)] = {"ms": ms, "value": S * world * n_live / (ms * 1e-3), "windows_per_s": S * world * (n_live / h) / (ms * 1e-3),
                                              "path_len_first_streams": lens}
        b.close()
    out["value"] = out["shapes"]["W100_h50"]["value"]
    return out if rank == 0 else None


# ----------------------------------------------------------------------------- striped single pair (cfg[4])
def bench_striped(ctx):
    """One long pair, column-striped over the ranks (K4): 25 000 columns per GPU (200k x 200k at 8 GPUs).
    Boundary columns travel over NVLink as peer stores from inside the stripe kernel; a step = accumulate on
    all ranks (CUDA events around each rank's launches, max over ranks)."""
    args, rank, world, torch, g, dist = ctx["args"], ctx["rank"], ctx["world"], ctx["torch"], ctx["g"], ctx["dist"]
    striped = g.submodule("striped")
    n = args.striped_cols_per_gpu * world
    live, ref = synth_chroma_pairs(1, n, 5000)
    a, b = np.ascontiguousarray(live[0]), np.ascontiguousarray(ref[0])
    sd = striped.StripedDtwDistributed(n, n, dist)
    c0, c1 = sd.bounds[rank]
    d_a = torch.from_numpy(a).cuda()
    d_b = torch.from_numpy(np.ascontiguousarray(b[:, c0:c1])).cuda()
    times = []
    for it in range(1 + max(2, min(args.steps, 3))):
        sd.reset()                       # ends with a barrier: all ranks launch together
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sd.accumulate(d_a, d_b)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        if it > 0:
            times.append(e0.elapsed_time(e1) * 1e-3)       # device time; the last stripe's kernel spans the whole critical path
    sd.backtrack()                       # first call also sets up NCCL's point-to-point channels
    dist.barrier()
    t0 = time.perf_counter()
    path = sd.backtrack()
    t_bt = time.perf_counter() - t0
    end = torch.tensor([sd.acc_end() if rank == world - 1 else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(end)
    t_acc = ctx["max_over_ranks"](float(np.mean(times)))
    out = None
    if rank == 0:
        d = np.diff(path, axis=0)
        ok = bool(path[0].tolist() == [0, 0] and path[-1].tolist() == [n - 1, n - 1] and ((d >= 0) & (d <= 1)).all() and (d.sum(axis=1) >= 1).all())
        # the same pair un-striped on this one GPU (K2 + K3; the direction map of n x n cells is n^2 / 4 bytes): the striped
        # result must be bit-equal — accumulated cost at the end cell and every path point — and it is the time to beat
        single = {"parity_vs_single_gpu": None, "single_gpu_ms": None}
        try:
            dtw = g.submodule("dtw")
            free_b = torch.cuda.mem_get_info()[0]
            if float(n) * n / 4 * 1.15 < free_b:
                plan1 = dtw.DtwPlan([n], [n], dtype="fp64")
                d_b_full = torch.from_numpy(b).cuda()
                plan1.accumulate(d_a, d_b_full)
                plan1.backtrack()
                torch.cuda.synchronize()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                plan1.accumulate(d_a, d_b_full)
                s1.record()
                plan1.backtrack()
                torch.cuda.synchronize()
                end1 = float(plan1.acc_end.cpu()[0])
                st1, ln1 = int(plan1.path_start.cpu()[0]), int(plan1.path_len.cpu()[0])
                path1 = plan1.path[st1 : st1 + ln1].cpu().numpy().astype(np.int64)
                same = bool(end1 == float(end.item()) and path1.shape == path.shape and np.array_equal(path1, path))
                single = {"parity_vs_single_gpu": same, "single_gpu_ms": s0.elapsed_time(s1), "single_gpu_acc_end": end1}
                plan1.close()
                del d_b_full
            else:
                single["note"] = "pair too large for one GPU's direction map"
        except Exception as exc:
            single["error"] = repr(exc)[:200]
        out = {"metric": "striped_dtw_gcups", "value": float(n) * n / t_acc / 1e9, "unit": "GCUPS", "n_gpus": world,
               "ms_per_step": t_acc * 1e3, "backtrack_ms": t_bt * 1e3, "scaling": "weak (25k columns per GPU)", "dtype": "f64",
               "config": {"workload": "single %d x %d pair column-striped over %d GPUs, NVLink peer-store boundary hand-off (BASELINE cfg[4] at 8 GPUs)" % (n, n, world)},
               "acc_end": float(end.item()), "path_len": int(len(path)), "path_valid": ok,
               "note": "a single pair is bound by the wavefront's critical path (bands x 64-step lag + columns), not by aggregate throughput"}
        out.update(single)
    sd.close()
    return out


def load_traffic(kernel, config):
    """DRAM bytes per launch of `kernel` at `config` from profiles/traffic.json (written by tools/make_traffic_json.py from
    an `ncu --set full` capture).  An entry records the SHA-1 of the kernel's source files at capture time: if a source
    has changed since, the figure is stale and None is reported instead."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            entries = json.load(fh)["entries"]
    except Exception:
        return None, "profiles/traffic.json missing"
    for e in entries:
        if e.get("kernel") == kernel and e.get("config") == config:
            for rel, sha in e.get("src_sha1", {}).items():
                try:
                    with open(os.path.join(ROOT, rel), "rb") as fh:
                        now = hashlib.sha1(fh.read()).hexdigest()
                except OSError:
                    now = None
                if now != sha:
                    return None, "stale: %s changed since the capture %s" % (rel, e.get("source"))
            return float(e["dram_bytes"]), e.get("source")
    return None, "no capture for this kernel / configuration"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def bench_dtw(ctx):
    args, rank, world, torch, nat, g = ctx["args"], ctx["rank"], ctx["world"], ctx["torch"], ctx["nat"], ctx["g"]
    barrier, max_over_ranks = ctx["barrier"], ctx["max_over_ranks"]
    dtw = g.submodule("dtw")
    # BASELINE cfg[2] as worded: ONE batch of 256 pairs, sharded over the ranks (strong scaling: 256 / N pairs per rank)
    P_total, Ln = args.pairs, args.length
    P = max(1, P_total // world)
    live, ref = synth_chroma_pairs(P, Ln, 2000 + rank * P)
    npdt = np.float64 if args.dtype == "fp64" else np.float32
    h_a = torch.from_numpy(np.ascontiguousarray(live, dtype=npdt)).pin_memory()
    h_b = torch.from_numpy(np.ascontiguousarray(ref, dtype=npdt)).pin_memory()
    plan = dtw.DtwPlan([Ln] * P, [Ln] * P, dtype=args.dtype)
    d_a = h_a.to("cuda")
    d_b = h_b.to("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    for _ in range(args.warmup_effective):
        plan.accumulate(d_a, d_b)
        plan.backtrack()
    barrier()
    launches0 = nat.launch_count()
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
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
    cells_rank = float(plan.cells)
    step_mode = "one step at a time on one stream (256 MiB L2 flush between steps)"
    if 2 * P <= DTW_PAIRS_PER_GPU:
        # A rank's share is a fraction of the batch: a launch of few pairs spends ~6 ms of its ~20 filling and draining the
        # chain of bands, so consecutive steps rotate over three plans on three streams and the next steps' bands fill the
        # SMs the previous one vacates (measured at 32 pairs per GPU: 24.5 ms per step alone, 24.3 with two in flight, 22.3
        # with three, no further gain with four).  Every step is still a full accumulate + backtrack of the rank's pairs; the timed
        # region is K steps between one start event and the later of the two streams' end events.
        depth = max(2, env_int("AFS_BENCH_DTW_DEPTH", 3))
        plans = [plan] + [dtw.DtwPlan([Ln] * P, [Ln] * P, dtype=args.dtype) for _ in range(depth - 1)]
        streams = [torch.cuda.Stream() for _ in range(depth)]

        def run_steps(k_steps):
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for st in streams:
                st.wait_event(e0)
            for k in range(k_steps):
                with torch.cuda.stream(streams[k % depth]):
                    plans[k % depth].accumulate(d_a, d_b)
                    plans[k % depth].backtrack()
            ends = []
            for st in streams:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record(st)
                ends.append(e1)
            torch.cuda.synchronize()
            return max(e0.elapsed_time(e1) for e1 in ends)

        run_steps(2)
        barrier()
        launches0 = nat.launch_count()
        total_ms = run_steps(args.steps)
        barrier()
        launches = nat.launch_count() - launches0
        step_ms = total_ms / args.steps
        step_mode = ("steps rotate over %d plans on %d streams (the next step's bands fill the SMs the previous step "
                     "vacates); no L2 flush: a step writes %.1f GB of direction map, far more than the 126 MB L2" % (depth, depth, cells_rank / 4 / 1e9))
        for pb in plans[1:]:
            pb.close()
        del plans
    step_ms_max = max_over_ranks(step_ms)
    value = cells_rank * world / (step_ms_max * 1e-3) / 1e9

    # ---- e2e: host (pinned) buffers in, paths out, through the plan object the drop-in API uses ----
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
    e2e_serial_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    # what limits e2e when all ranks upload at once: this rank's host -> device rate with every rank copying concurrently
    # (CUDA events around the two uploads alone, three repetitions, slowest rank reported)
    up0, up1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    up0.record()
    for _ in range(3):
        da = h_a.to("cuda", non_blocking=True)
        db = h_b.to("cuda", non_blocking=True)
    up1.record()
    torch.cuda.synchronize()
    up_bytes = 3 * (h_a.numel() * h_a.element_size() + h_b.numel() * h_b.element_size())
    h2d_gbs_slowest = up_bytes / max_over_ranks(up0.elapsed_time(up1) * 1e-3) / 1e9
    del da, db

    # The same work as a stream of batches (what a service does), through the public `dtw.DtwPipeline`: two batches are in
    # flight on two plans / two compute streams, so the ramp-down of one launch and its backtrack overlap the ramp-up of
    # the next; a batch's inputs are uploaded on a copy stream while the previous batch computes, and its paths are read
    # back while the next one computes.  Every step still moves its own inputs host -> device and its own results
    # device -> host inside the timed region.
    pipe_depth = env_int("AFS_BENCH_PIPE_DEPTH", 3)     # three batches in flight (measured at 256 pairs: 609 GCUPS with two, 621 with three)
    pipe = dtw.DtwPipeline([Ln] * P, [Ln] * P, dtype=args.dtype, depth=pipe_depth)

    def run_pipelined(n):
        for _ in range(n):
            pipe.submit(h_a, h_b, as_arrays=False)
        pipe.drain(as_arrays=False)
        torch.cuda.synchronize()

    run_pipelined(2)
    barrier()
    pipe_steps = 2 * e2e_steps                    # the pipeline's fill and drain are inside the timed region
    t0 = time.perf_counter()
    run_pipelined(pipe_steps)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / pipe_steps)
    e2e_val = cells_rank * world / e2e_s / 1e9
    e2e_serial_val = cells_rank * world / e2e_serial_s / 1e9
    pipe.close()
    del pipe
    h2d = int(h_a.numel() * h_a.element_size() + h_b.numel() * h_b.element_size())
    d2h = int(h_path.numel() * 4 + h_start.numel() * 4 + h_len.numel() * 4 + h_end.numel() * 8)
    # ---- weak-scaling companion (N > 1): the full batch on every GPU, as round 1 reported it ----
    weak = None
    if world > 1 and P < P_total and P_total % P == 0:
        rep = P_total // P
        wa, wb = d_a.repeat(rep, 1, 1), d_b.repeat(rep, 1, 1)
        wplan = dtw.DtwPlan([Ln] * P_total, [Ln] * P_total, dtype=args.dtype)
        wplan.accumulate(wa, wb)
        wplan.backtrack()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(2):
            wplan.accumulate(wa, wb)
            wplan.backtrack()
        w1.record()
        barrier()
        wms = max_over_ranks(w0.elapsed_time(w1) / 2)
        weak = {"value": float(wplan.cells) * world / (wms * 1e-3) / 1e9, "unit": "GCUPS", "scaling": "weak", "pairs_per_gpu": P_total,
                "ms_per_step": wms, "steps": 2, "note": "%d pairs on every GPU (the rank's %d pairs repeated)" % (P_total, P)}
        wplan.close()
        del wplan, wa, wb
        torch.cuda.empty_cache()
    # ---- the other arithmetic mode on the same data (kernel only): fp32 offsets against fp64 bases ----
    other = None
    if args.dtype == "fp64":
        end64 = plan.acc_end.cpu().numpy().copy()
        plan32 = dtw.DtwPlan([Ln] * P, [Ln] * P, dtype="fp32")
        a32, b32 = d_a.to(torch.float32), d_b.to(torch.float32)
        for _ in range(2):
            plan32.accumulate(a32, b32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan32.accumulate(a32, b32)
        e1.record()
        torch.cuda.synchronize()
        ms32 = max_over_ranks(e0.elapsed_time(e1) / 3)
        end32 = plan32.acc_end.cpu().numpy()
        other = {"dtype": "f32 (re-based offsets, fp64 bases)", "accumulate_ms": ms32,
                 "value": cells_rank * world / (ms32 * 1e-3) / 1e9, "unit": "GCUPS (accumulate kernel only)",
                 "max_rel_err_acc_end_vs_f64": float(np.max(np.abs(end32 - end64) / np.abs(end64))), "tolerance": 1e-5}
        plan32.close()
        del a32, b32
    if rank != 0:
        return None
    peaks = load_peaks()
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    lanes = FP64_LANES_PER_SM if args.dtype == "fp64" else 128
    pipe_peak_tflops = n_sm * lanes * 2 * sm_max * 1e6 / 1e12
    acc_ms_mean = float(np.mean(acc_ms))
    achieved_tflops = cells_rank * DTW_FLOP_PER_CELL / (acc_ms_mean * 1e-3) / 1e12
    alg_bytes = cells_rank * 0.25 + 2 * P * 12 * Ln * (8 if args.dtype == "fp64" else 4)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    dtw_traffic = load_traffic("dtw_wavefront_kernel<%s>" % ("double" if args.dtype == "fp64" else "float"), "%dx%d" % (P, Ln))
    roofline = {
        "kernel": "dtw_wavefront_kernel<%s>" % ("double" if args.dtype == "fp64" else "float"),
        "bound": "fp64_pipe" if args.dtype == "fp64" else "fp32_pipe",
        "achieved": achieved_tflops, "peak": pipe_peak_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / pipe_peak_tflops,
        "peak_source": "derived: %d SMs x %d lanes x 2 x %.0f MHz (MEASURED_PEAKS.json has no %s pipe figure)" % (n_sm, lanes, sm_max, args.dtype),
        "flop_per_cell": DTW_FLOP_PER_CELL, "kernel_ms": acc_ms_mean,
        "practical_ceiling": "tools/ubench_split.cu on B200: the step's arithmetic alone (cost FMAs + DP chain + shuffle, no hand-off / TMA / "
                             "stores) takes 225 cycles per 128-cell step per sub-partition = ~660 GCUPS (fp64); DFMA with 3 distinct register "
                             "operands costs 2.6 pipe cycles (nominal 2.0), see DESIGN.md K2",
        "hbm": {"achieved": alg_bytes / (acc_ms_mean * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (acc_ms_mean * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "measured" if "hbm_gbs" in peaks else "fallback"},
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch at exactly this configuration; None for any other.
        # 256 pairs: the packed seq_b of all pairs (573 MB) no longer fits L2, so the per-band re-reads of seq_b reach DRAM
        # (80.7 GB read + 37.3 GB written; 0.76 TB/s = 12 % of the HBM peak, not limiting: see DESIGN.md K2).
        "traffic": dtw_traffic[0], "traffic_source": dtw_traffic[1],
        "algorithmic_bytes": alg_bytes,
    }
    cpu = None if args.no_cpu_baseline else cpu_dtw_baseline(seconds=args.cpu_seconds)
    return {
        "metric": "dtw_gcups", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup_effective, "ms_per_step": step_ms_max, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64" if args.dtype == "fp64" else "f32", "data": "synthetic",
        "config": {"workload": "offline full DTW %dx%d chroma frames, one batch of %d pairs sharded over %d GPU(s): %d pairs per GPU (BASELINE cfg[2]; strong scaling)" % (Ln, Ln, P * world, world, P),
                   "pairs_total": P * world, "pairs_per_gpu": P, "frames": Ln, "features": 12,
                   "l2": "per-step working set %.1f GB of direction map per GPU > 126 MB L2" % (cells_rank / 4 / 1e9),
                   "step": "accumulate (K2) + backtrack (K3)", "step_mode": step_mode,
                   "parallelism": "pairs sharded over ranks, no collective"},
        "kernel_ms": {"accumulate": acc_ms_mean, "backtrack": float(np.mean(bt_ms))},
        "fp32_mode": other, "weak": weak,
        "wall_s_timed_region": t_wall,
        "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": 2 * e2e_steps,
                "mode": "dtw.DtwPipeline(depth=%d): a stream of batches, that many in flight on their own plans / compute streams; a batch's " % pipe_depth +
                        "upload and read-back overlap the neighbouring batch's kernels (copy stream + events); every step's own "
                        "copies are inside the timed region",
                "h2d_gbs_per_rank_all_ranks_copying": h2d_gbs_slowest,
                "serial_value": e2e_serial_val,
                "serial_note": "one batch at a time: upload, K2, K3, read-back, synchronize"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }


# ----------------------------------------------------------------------------- chroma (cfg[1])
CHROMA_TRACKS = 1024
CHROMA_SECONDS = 300.0
CHROMA_BYTES_PER_FRAME = 2048 * 4 + 12 * 4      # hop samples read once + 12 floats out
CHROMA_FLOP_PER_FRAME = 1.9e5


def synth_audio_tracks(torch, n_tracks, n_samples, seed, device):
    """SURVEY.md §8(d) cfg2 generator on the device: random note sequence (0.5 s notes, MIDI 36-84,
    6 harmonics with 1/h amplitude) + 0.01 N(0,1) noise, float32 in [-1, 1]."""
    gen = torch.Generator(device=device).manual_seed(seed)
    note = 11025
    n_notes = (n_samples + note - 1) // note
    out = torch.empty((n_tracks, n_samples), dtype=torch.float32, device=device)
    ph = torch.arange(note, device=device, dtype=torch.float32) * (2 * np.pi / 22050.0)
    chunk = 16
    for t0 in range(0, n_tracks, chunk):
        t1 = min(n_tracks, t0 + chunk)
        midi = torch.randint(36, 85, (t1 - t0, n_notes), generator=gen, device=device).to(torch.float32)
        freq = 440.0 * torch.pow(2.0, (midi - 69.0) / 12.0)
        phase = freq[:, :, None] * ph[None, None, :]                       # (tracks, notes, note)
        x = torch.zeros_like(phase)
        for hmn in range(1, 7):
            x += torch.sin(hmn * phase) / hmn
        x = (0.25 * x).reshape(t1 - t0, -1)[:, :n_samples]
        x += 0.01 * torch.randn(x.shape, generator=gen, device=device)
        out[t0:t1] = x.clamp_(-1.0, 1.0)
    return out


def cpu_chroma_baseline(seconds=8.0, threads=None):
    """numpy oracle (chroma.py:44-75 restated with numpy's own rfft) on a 30 s synthetic track per call."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import afs_oracle as orc
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(1000)
    x = (0.2 * np.sin(2 * np.pi * 440.0 * np.arange(30 * 22050) / 22050.0) + 0.01 * rng.standard_normal(30 * 22050)).astype(np.float32)
    t_end = time.perf_counter() + seconds
    counts = [0] * threads

    def work(k):
        while time.perf_counter() < t_end:
            counts[k] += orc.wav_samples_to_chroma(x).shape[1]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    return {"value": sum(counts) / dt, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": "%d frames of 30 s synthetic tracks (numpy rfft oracle) over %d threads in %.1f s" % (sum(counts), threads, dt)}


def bench_chroma(ctx):
    args, rank, world, torch, nat, g = ctx["args"], ctx["rank"], ctx["world"], ctx["torch"], ctx["nat"], ctx["g"]
    barrier, max_over_ranks = ctx["barrier"], ctx["max_over_ranks"]
    ch = g.submodule("chroma")
    T = args.chroma_tracks
    n = int(CHROMA_SECONDS * 22050)            # 6 615 000 samples
    plan = ch.default_plan()
    compute = args.chroma_compute          # "tc": tcgen05 DFT-as-GEMM pipeline (default); "fp32": CUDA-core FFT kernel
    audio = synth_audio_tracks(torch, T, n, 1000 + rank, "cuda").reshape(-1)
    offs = np.arange(T + 1, dtype=np.int64) * n
    out, foffs = plan.run(audio, offs, compute=compute)
    frames = int(foffs[-1])
    steps = args.steps
    for _ in range(args.warmup_effective):
        plan.run(audio, offs, d_out=out, compute=compute)
    barrier()
    launches0 = nat.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for k in range(steps):
        ev[k][0].record()
        plan.run(audio, offs, d_out=out, compute=compute)          # input 27 GB per step >> 126 MB L2: no flush needed
        ev[k][1].record()
    barrier()
    launches = nat.launch_count() - launches0
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    ms_max = max_over_ranks(ms)
    value = frames * world / (ms_max * 1e-3)
    # ---- e2e: pinned host audio -> H2D in slabs -> K1 -> D2H chroma; 64 distinct host tracks reused for all slots ----
    slab_tracks = 64
    h_slab = audio[: slab_tracks * n].cpu().pin_memory()
    slab_offs = np.arange(slab_tracks + 1, dtype=np.int64) * n
    d_slabs = [torch.empty(slab_tracks * n, dtype=torch.float32, device="cuda") for _ in range(2)]
    frames_slab = frames // T * slab_tracks
    d_outs = [torch.empty(12 * frames_slab, dtype=torch.float32, device="cuda") for _ in range(2)]
    h_out = torch.empty((T // slab_tracks, 12 * frames_slab), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()

    def step_e2e(h_src, d_bufs):
        evs = []
        for i in range(T // slab_tracks):
            b = i & 1
            if i >= 2:
                copy_stream.wait_event(evs[i - 2])      # slab b is free once its previous kernel + D2H are done
            with torch.cuda.stream(copy_stream):
                d_bufs[b].copy_(h_src, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            torch.cuda.current_stream().wait_event(ready)
            plan.run(d_bufs[b], slab_offs, d_out=d_outs[b], compute=compute)
            h_out[i].copy_(d_outs[b], non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            evs.append(done)
        torch.cuda.synchronize()

    def time_e2e(h_src, d_bufs, sample_bytes, what):
        step_e2e(h_src, d_bufs)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = 2
        for _ in range(e2e_steps):
            step_e2e(h_src, d_bufs)
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        return {"value": frames * world / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(T * n * sample_bytes),
                "d2h_bytes_per_step": int(12 * frames * 4), "steps": e2e_steps,
                "note": "%s; 64 distinct pinned host tracks reused to fill all %d track slots; double-buffered slabs" % (what, T)}

    e2e = e2e_pcm = pcm_resident = None
    if T % slab_tracks == 0:
        e2e = time_e2e(h_slab, d_slabs, 4, "float32 samples on the host (what librosa.load returns)")
        # the same tracks as 16-bit PCM (what the WAV files hold): half the PCIe bytes, 1/32768 applied in the kernel
        del d_slabs
        h_pcm = (h_slab * 32768.0).round().clamp_(-32768, 32767).to(torch.int16).pin_memory()
        d_pcm = [torch.empty(slab_tracks * n, dtype=torch.int16, device="cuda") for _ in range(2)]
        e2e_pcm = time_e2e(h_pcm, d_pcm, 2, "int16 PCM samples on the host (afs_chroma_batch_pcm16)")
        plan.run(d_pcm[0], slab_offs, d_out=d_outs[0], compute=compute)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(3):
            plan.run(d_pcm[0], slab_offs, d_out=d_outs[0], compute=compute)
        p1.record()
        torch.cuda.synchronize()
        pcm_resident = {"value": frames_slab * 3 / (p0.elapsed_time(p1) * 1e-3), "unit": "frames/s per GPU",
                        "note": "int16 samples resident in HBM, %d tracks per launch" % slab_tracks}
    if rank != 0:
        return None
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
    gbs = frames * CHROMA_BYTES_PER_FRAME / (ms * 1e-3) / 1e9
    tfl = frames * CHROMA_FLOP_PER_FRAME / (ms * 1e-3) / 1e12
    cpu = None if args.no_cpu_baseline else cpu_chroma_baseline(seconds=min(8.0, args.cpu_seconds))
    kernel_name = "chroma_tc_spectrum_kernel" if compute == "tc" else "chroma_fast_kernel<17>"
    ch_tr = load_traffic(kernel_name, "per frame")
    return {
        "metric": "chroma_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup_effective, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
        "data": "synthetic", "vs_baseline": None,
        "config": {"workload": "batched chroma extraction: %d synthetic 5-min 22.05 kHz tracks per GPU, n_fft=4096, hop=2048 (BASELINE cfg[1])" % T,
                   "tracks_per_gpu": T, "frames_per_step": frames, "l2": "27 GB of input per step >> 126 MB L2"},
        "e2e": e2e, "e2e_pcm16": e2e_pcm, "pcm16_resident": pcm_resident, "gpu_launches": int(launches),
        "compute": compute,
        "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": gbs / hbm_peak, "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                     "bytes_per_frame": CHROMA_BYTES_PER_FRAME, "kernel_ms": ms,
                     "traffic": None if ch_tr[0] is None else ch_tr[0] * frames, "traffic_source": ch_tr[1],
                     "traffic_note": "capture of a smaller launch of the same kernel, DRAM bytes per frame x frames of this launch",
                     "fp32": {"achieved": tfl, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tfl / fp32_peak,
                              "flop_per_frame": CHROMA_FLOP_PER_FRAME,
                              "peak_source": "derived: %d SMs x 128 lanes x 2 x %.0f MHz" % (n_sm, sm_max)}},
        "cpu_baseline": cpu,
    }


# ----------------------------------------------------------------------------- OTW latency (cfg[3])
OTW_STREAMS = 4096
OTW_REF = 3000
OTW_LIVE = 3300
OTW_C = 500


def synth_streams(torch, n_streams, n_ref, n_live, seed, device):
    """per-stream AR(1)-smoothed random chroma reference, live = warped copy + noise."""
    gen = torch.Generator(device=device).manual_seed(seed)
    x = torch.rand((n_streams * 12, 1, n_ref + 63), generator=gen, device=device, dtype=torch.float64)
    taps = (0.3 * 0.7 ** torch.arange(63, -1, -1, device=device, dtype=torch.float64)).reshape(1, 1, 64)
    y = torch.nn.functional.conv1d(x, taps).reshape(n_streams, 12, n_ref)
    ref = y / y.norm(dim=1, keepdim=True)
    u = torch.linspace(0, 1, n_live, device=device, dtype=torch.float64)
    idx = torch.round(torch.clamp(u + 0.08 * torch.sin(6 * np.pi * u), 0, 1) * (n_ref - 1)).long()
    live = ref[:, :, idx] + 0.05 * torch.rand((n_streams, 12, n_live), generator=gen, device=device, dtype=torch.float64)
    live = live / live.norm(dim=1, keepdim=True)
    frames = live.permute(2, 0, 1).contiguous()            # (n_live, n_streams, 12)
    return ref.contiguous(), frames


def cpu_otw_baseline(frames_per_thread=400, threads=None):
    """Oracle C port of OnlineTimeWarping.insert (otw_eran.py:38-85), c = 500, one stream per host thread."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import afs_oracle as orc
    threads = threads or os.cpu_count() or 1
    live, ref = synth_chroma_pairs(1, 1200, 3000)
    live, ref = live[0], ref[0]
    lat = [[] for _ in range(threads)]

    def work(k):
        o = orc.OnlineTimeWarping(ref, {"c": OTW_C, "max_run_count": 3})
        for i in range(min(live.shape[1], 600 + frames_per_thread)):
            t0 = time.perf_counter()
            r = o.insert(live[:, i])
            if i >= 600:
                lat[k].append(time.perf_counter() - t0)
            if r == "stop":
                break

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    allv = np.array([v for l in lat for v in l]) * 1e3
    return {"value": float(np.percentile(allv, 99)), "unit": "ms p99 per frame per stream", "p50_ms": float(np.percentile(allv, 50)),
            "cores": threads, "kind": "port",
            "streams_per_ms_all_cores": threads / float(np.mean(allv)),
            "sample": "%d inserts (steady state, c=500) on %d threads, one stream each, %.1f s" % (len(allv), threads, dt)}


def bench_otw(ctx):
    args, rank, world, torch, nat, g = ctx["args"], ctx["rank"], ctx["world"], ctx["torch"], ctx["nat"], ctx["g"]
    barrier, max_over_ranks = ctx["barrier"], ctx["max_over_ranks"]
    batch = g.submodule("batch")
    S = args.otw_streams
    ref, frames = synth_streams(torch, S, OTW_REF, OTW_LIVE, 3000 + rank, "cuda")
    res = {}
    for kind in ("otw", "livenote_v2"):
        b = batch.OtwBatch(ref, OTW_C, 3, kind=kind)
        st, npts, pts = b._outputs(1)
        d_flat = b.outputs_flat(1)                      # status | npoints | points: one buffer, one D2H copy per step
        h_flat = torch.empty_like(d_flat, device="cpu").pin_memory()
        h_frames = frames.cpu().pin_memory() if kind == "otw" else None
        warm = OTW_C + 100
        n_meas = min(2000, OTW_LIVE - warm - 100) if args.otw_steps <= 0 else args.otw_steps
        lat = []
        launches0 = nat.launch_count()
        dev_ms = []
        for k in range(warm + n_meas):
            fr = frames[k]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            b.step_device(fr)
            e1.record()
            h_flat.copy_(d_flat, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if k >= warm:
                lat.append(dt * 1e3)
                dev_ms.append(e0.elapsed_time(e1))
        launches = nat.launch_count() - launches0
        # e2e: the frame batch starts in pinned HOST memory (H2D inside the timed region)
        lat_e2e = []
        if h_frames is not None:
            b.reset()
            d_fr = torch.empty_like(frames[0])
            for k in range(warm + 500):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                d_fr.copy_(h_frames[k], non_blocking=True)
                b.step_device(d_fr)
                h_flat.copy_(d_flat, non_blocking=True)
                torch.cuda.synchronize()
                if k >= warm:
                    lat_e2e.append((time.perf_counter() - t0) * 1e3)
        lat = np.array(lat)
        p99 = max_over_ranks(float(np.percentile(lat, 99)))
        p50 = max_over_ranks(float(np.percentile(lat, 50)))
        res[kind] = {"p99_ms": p99, "p50_ms": p50, "max_ms": float(lat.max()), "steps": int(len(lat)),
                     "kernel_ms_mean": float(np.mean(dev_ms)), "kernel_ms_p99": float(np.percentile(dev_ms, 99)),
                     "gpu_launches": int(launches),
                     # per-step latency histogram of this rank (the evidence asked for K5)
                     "latency_percentiles_ms": {q: float(np.percentile(lat, float(q))) for q in ("10", "50", "90", "99", "99.9")},
                     "latency_histogram": {"edges_ms": [round(float(e), 4) for e in np.histogram_bin_edges(lat, bins=12)],
                                           "counts": [int(c) for c in np.histogram(lat, bins=12)[0]]},
                     "stream_steps_per_s": S * world / (float(np.mean(lat)) * 1e-3)}
        if lat_e2e:
            le = np.array(lat_e2e)
            res[kind]["e2e"] = {"value": max_over_ranks(float(np.percentile(le, 99))), "unit": "ms p99 per frame", "p50_ms": float(np.percentile(le, 50)),
                                "h2d_bytes_per_step": int(S * 12 * 8), "d2h_bytes_per_step": int(h_flat.numel() * 4),
                                "steps": int(len(le))}
        b.close()
    if rank != 0:
        return None
    cpu = None if args.no_cpu_baseline else cpu_otw_baseline()
    o = res["otw"]
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # The reference and live windows (c x 96 B each) are re-read every step and do not fit L2 for 4096 streams.
    # ncu --set full of ONE heavy step (t = 703, all 4096 streams run a row sweep and a column sweep): 1.03 GB read +
    # 0.06 GB written.  It is set against the live p99 kernel time (the heavy steps), not the mean.
    # Algorithmic bytes of a heavy step: every stream re-reads its reference and live windows, c columns x 96 B each
    # (2 x 500 x 96 B x 4096 streams = 0.39 GB), and they do not fit L2.  Set against the live p99 kernel time (the heavy steps).
    alg = 2.0 * OTW_C * 96 * S
    heavy_ms = o["kernel_ms_p99"]
    tr, tr_src = load_traffic("otw_step_kernel", "%d streams c=%d heavy step" % (S, OTW_C))
    otw_roof = {"kernel": "otw_step_kernel", "bound": "hbm", "achieved": alg / (heavy_ms * 1e-3) / 1e9,
                "peak": hbm_peak, "unit": "GB/s", "frac": alg / (heavy_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "measured" if "hbm_gbs" in peaks else "fallback", "algorithmic_bytes": alg, "traffic": tr, "kernel_ms": heavy_ms,
                "traffic_source": tr_src,
                "note": "the metric is latency; the step is a chain of <= 5 serial sweeps per stream on top of this traffic"}
    return {
        "metric": "otw_p99_frame_latency_ms", "value": o["p99_ms"], "unit": "ms", "n_gpus": world, "higher_is_better": False,
        "steps": o["steps"], "warmup": OTW_C + 100, "ms_per_step": o["p50_ms"], "scaling": "weak", "dtype": "f64", "data": "synthetic", "vs_baseline": None,
        "config": {"workload": "online OTW (otw_eran) streaming alignment, %d concurrent streams per GPU, c=%d, max_run_count=3, ref %d frames (BASELINE cfg[3])" % (S, OTW_C, OTW_REF),
                   "latency": "host wall time from 'frame batch resident in HBM' to 'status/points of all streams visible on the host', one launch per frame"},
        "otw": o, "livenote_v2": res["livenote_v2"], "e2e": o.get("e2e"), "gpu_launches": o["gpu_launches"],
        "roofline": otw_roof, "cpu_baseline": cpu,
    }


if __name__ == "__main__":
    main()
