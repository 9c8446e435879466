#!/usr/bin/env python
"""Benchmark of the hot path: normals + RSD points/s at r = 2 cm on the 20 M-point synthetic room
cloud (BASELINE.json config C4), query-sharded over N B200s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference

One "step" = one pass of the hot path over the cloud: grid build (replaces the kd-tree build) +
normals + RSD.  `value` is measured with the cloud resident in HBM; `e2e` goes through the C ABI
with pinned host buffers (H2D of the cloud, D2H of normals and radii inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "normals+RSD points/s (r=2cm)"
UNIT = "points/s"
RADIUS = 0.02
NDIV = 10
PLANE_RADIUS = 0.1
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=20_000_000, help="cloud size (default: config C4)")
    ap.add_argument("--exact", action="store_true", help="fp64 accumulation mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=6_000_000)
    ap.add_argument("--workload", default="normals_rsd", choices=["normals_rsd", "grsd"],
                    help="normals_rsd: the headline C4 metric; grsd: config C3, 512 clusters, GRSD clouds/s")
    ap.add_argument("--clusters", type=int, default=512)
    ap.add_argument("--gather", action="store_true", help="N>1: also all-gather the results into every rank's HBM")
    return ap.parse_args()


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def slab_sample(pts, target):
    """A contiguous x-slab of the cloud holding ~target points (keeps the surface density)."""
    import numpy as np

    if pts.shape[0] <= target:
        return pts
    xs = np.sort(pts[:: max(1, pts.shape[0] // 200_000), 0])
    frac = target / pts.shape[0]
    lo = xs[int(0.35 * len(xs))]
    hi = xs[min(len(xs) - 1, int((0.35 + frac) * len(xs)))]
    return np.ascontiguousarray(pts[(pts[:, 0] >= lo) & (pts[:, 0] < hi)])


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML (what nvidia-smi reads) is polled from a
    thread every few milliseconds, so that even a 50 ms region (8 GPUs) gets samples; nvidia-smi -lms is the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_s=0.004):
        self.index = index
        self.period_s = period_s
        self.proc = None
        self.thread = None
        self.rows = []  # (sm MHz, max MHz, watts, reasons bit mask)
        self._stop = False
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample(self):
        nv = self.nvml
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            try:
                watts = nv.nvmlDeviceGetPowerUsage(self.handle) / 1e3
            except Exception:
                watts = None
            self.rows.append((float(sm), float(mx), watts, int(mask)))
        except Exception:
            pass

    def _poll(self):
        while not self._stop:
            self._sample()
            time.sleep(self.period_s)

    def start(self):
        if self.nvml:
            import threading

            self._stop = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nvml:
            self._stop = True
            if self.thread:
                self.thread.join(timeout=2)
            nv = self.nvml
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            sm = [r[0] for r in self.rows]
            mx = [r[1] for r in self.rows]
            power = [r[2] for r in self.rows if r[2] is not None]
            reasons = sorted(k for k, b in bits.items() if any(r[3] & b for r in self.rows))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                    "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def cpu_points_per_s(pts, nthreads=0):
    """Oracle (CPU restatement) normals + RSD on the given sample, all host threads."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import pyoracle

    pyoracle.build()
    t0 = time.perf_counter()
    n4, k = pyoracle.normals(pts, RADIUS, nthreads=nthreads)
    rmin, rmax, _ = pyoracle.rsd(pts, n4, RADIUS, ndiv=NDIV, plane_radius=PLANE_RADIUS, nthreads=nthreads)
    dt = time.perf_counter() - t0
    cpu_points_per_s.last = (n4, k, rmin, rmax)  # kept for the parity gate of the same run
    return pts.shape[0] / dt, dt, pyoracle.num_threads() if nthreads <= 0 else nthreads


def c1_reference_faithful(ctx):
    """BASELINE config C1 (the reference's own CPU-runnable case: sample_pipeline's RSD on a 100 k-point tabletop cloud,
    r = 2 cm, the plugin's default max_nn = 150) the way the reference runs it -- kd-tree, materialised neighbour lists,
    estimation loop, ONE thread, the three phases timed where radius_estimation.cpp:108,125,216 log them -- next to the
    same call through the C ABI (upload + grid + thresholds + RSD + download, host buffers).  A reported baseline."""
    import numpy as np

    import pyoracle
    from mapping_private_b200 import synth

    pts = synth.tabletop(100_000)
    n4, _ = pyoracle.normals(pts, RADIUS)
    nrm = np.ascontiguousarray(n4[:, :3])
    t0 = time.perf_counter()
    omin, omax, phases = pyoracle.rsd_ref_faithful(pts, nrm, RADIUS, max_nn=150, ndiv=NDIV, plane_radius=PLANE_RADIUS)
    cpu_s = time.perf_counter() - t0
    ctx.upload(pts)  # warm-up of this size
    ctx.build_grid(RADIUS)
    ctx.set_normals(n4)
    ctx.rsd(RADIUS, max_nn=150, ndiv=NDIV, plane_radius=PLANE_RADIUS)
    t0 = time.perf_counter()
    ctx.upload(pts)
    ctx.build_grid(RADIUS)
    ctx.set_normals(n4)
    gmin, gmax = ctx.rsd(RADIUS, max_nn=150, ndiv=NDIV, plane_radius=PLANE_RADIUS)
    gpu_s = time.perf_counter() - t0
    rel = float(np.nanmax(np.maximum(np.abs(gmin - omin) / omin, np.abs(gmax - omax) / omax)))
    return {"workload": "C1 100k-point synthetic tabletop, RSD r=2cm, max_nn=150 (plugin default), normals given",
            "cpu_points_per_s": pts.shape[0] / cpu_s, "cpu_cores": 1, "cpu_kind": "port, reference-faithful organisation",
            "cpu_phases_s": {"kdtree_build": float(phases[0]), "neighbour_search": float(phases[1]), "estimation": float(phases[2])},
            "b200_points_per_s_e2e": pts.shape[0] / gpu_s, "b200_ms_e2e": 1e3 * gpu_s, "radii_max_rel_err": rel}


def parity_gate(ctx, sample):
    """the parity gates that go with every benchmark line.  The CUDA path on the slab the CPU baseline just
    processed, against that run's oracle results: neighbour counts bit-exact, normals within 1e-4 rad (sign-insensitive),
    radii within 1e-4 relative given the same normals.  The oracle is the checker here, nothing of it is timed."""
    import numpy as np

    o4, ok, omin, omax = cpu_points_per_s.last
    ctx.upload(sample)
    ctx.build_grid(RADIUS)
    g4 = ctx.normals(RADIUS)
    same_counts = int(ctx.profile()["neighbour_sum"]) == int(ok.sum())
    good = ~np.isnan(o4[:, 0]) & ~np.isnan(g4[:, 0])
    sin_angle = np.linalg.norm(np.cross(g4[good, :3].astype(np.float64), o4[good, :3].astype(np.float64)), axis=1)
    ctx.set_normals(o4)
    gmin, gmax = ctx.rsd(RADIUS, ndiv=NDIV, plane_radius=PLANE_RADIUS)
    rel = np.maximum(np.abs(gmin - omin) / omin, np.abs(gmax - omax) / omax)
    return {"sample_points": int(sample.shape[0]), "neighbour_counts_equal": bool(same_counts),
            "nan_normals_equal": bool(np.array_equal(np.isnan(o4[:, 0]), np.isnan(g4[:, 0]))),
            "normals_fraction_over_1e-4_rad": float(np.mean(sin_angle > 1e-4)), "normals_p99.9_rad": float(np.percentile(sin_angle, 99.9)),
            "radii_max_rel_err_given_same_normals": float(np.nanmax(rel)), "mode": "fast-fp32 kernels vs fp64 oracle"}


def run_reference(args, rank):
    """--impl reference: the reference's algorithm on the host cores (oracle port: the reference
    plugins need ROS/PCL/ANN, none of which exist here -- DESIGN.md "Oracle")."""
    if rank != 0:
        return
    import pkgpath

    pkgpath.load()
    from mapping_private_b200 import synth

    pts = synth.room(args.points)  # full cloud so that the slab has the workload's density
    sample = slab_sample(pts, 1_500_000)  # ~3-4 s of work per step on 16 cores: K + W steps stay within minutes
    rates = []
    cores = 1
    for s in range(args.warmup + args.steps):
        rate, dt, cores = cpu_points_per_s(sample)
        if s >= args.warmup:
            rates.append((rate, dt))
    value = len(rates) * sample.shape[0] / sum(dt for _, dt in rates)
    ms = 1e3 * sum(dt for _, dt in rates) / len(rates)
    desc = f"x-slab of the room cloud, {sample.shape[0]} points per step (same density as the {args.points}-point workload)"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 20M-point synthetic room, normals+RSD r=2cm (bounded sample per step)", "points": args.points,
                   "radius_m": RADIUS, "distance_div": NDIV, "plane_radius": PLANE_RADIUS, "max_nn": "unlimited"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_grsd(args, rank, world, local_rank):
    """Config C3: GRSD-21 of a batch of segmented clusters (2.5 cm voxels), cluster-per-GPU.
    Host buffers in, 21 int32 bins per cluster out (the GlobalRSD plugin's work); histograms of the
    ranks are summed with one all-reduce."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from mapping_private_b200 import cab, shard, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    xyz, off = synth.clusters(args.clusters)
    sizes = np.diff(off)
    mine = shard.assign_clusters_lpt(sizes.tolist(), world)[rank]
    my_xyz = np.concatenate([xyz[off[c]:off[c + 1]] for c in mine]) if mine else np.zeros((0, 3), np.float32)
    my_xyz = torch.from_numpy(np.ascontiguousarray(my_xyz)).pin_memory().numpy()  # page-locked host buffer (H2D at PCIe speed)
    my_off = np.concatenate([[0], np.cumsum(sizes[mine])]).astype(np.int32)
    ctx = cab.Context(local_rank, exact=True)
    leaf = 0.025

    def step():
        hist = torch.zeros((args.clusters, 21), dtype=torch.int32)
        if mine:
            h = ctx.grsd_batch(my_xyz, my_off, leaf, r_normals=0.02)
            hist[mine] = torch.from_numpy(h)
        if world > 1:
            hd = hist.to(dev)
            shard.allreduce_histograms(hd)
            hist = hd.cpu()
        return hist

    for _ in range(max(args.warmup, 3)):
        hist = step()
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ctx.profile()["kernel_launches"]
    sampler.start()
    t0 = time.perf_counter()
    kern_ms = []
    for _ in range(args.steps):
        hist = step()
        p = ctx.profile()
        kern_ms.append(p["build_ms"] + p["normals_ms"] + p["grsd_ms"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.profile()["kernel_launches"] - l0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, str(ROOT / "oracle"))
        import pyoracle

        pyoracle.build()
        ncheck = min(args.clusters, 48)
        t1 = time.perf_counter()
        bad = 0
        for c in range(ncheck):
            o = pyoracle.grsd21(xyz[off[c]:off[c + 1]], leaf, r_normals=0.02)
            bad += int(not np.array_equal(o["hist21"], hist[c].numpy()))
        cdt = time.perf_counter() - t1
        cpu = {"value": ncheck / cdt, "unit": "clouds/s", "cores": pyoracle.num_threads(), "kind": "port",
               "sample": f"first {ncheck} clusters, oracle normals+voxel RSD+GRSD; histograms differing from the GPU: {bad}"}
    if rank == 0:
        ms = 1e3 * dt / args.steps
        emit(json.dumps({
            "metric": "GRSD clouds/s", "value": args.clusters / (dt / args.steps), "unit": "clouds/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3 GRSD-21 on {args.clusters} synthetic clusters, leaf 2.5 cm, normals r=2cm, cluster-per-GPU",
                       "points": int(off[-1]), "leaf_m": leaf, "parallelism": f"clusters LPT x{world}, int32 all-reduce of histograms",
                       "timed": "pinned host buffers in, histograms out (H2D + D2H inside)"},
            "kernels_ms_per_step_rank0": statistics.mean(kern_ms), "cpu_baseline": cpu, "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": args.clusters / (dt / args.steps), "unit": "clouds/s", "h2d_bytes_per_step": int(off[-1]) * 12,
                    "d2h_bytes_per_step": args.clusters * 84},
        }))
    if world > 1:
        dist.destroy_process_group()


class _DevArray:
    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def bind_to_gpu_numa_node(index: int):
    """Pins this rank to the CPU cores next to its GPU (NVML's affinity mask) before any host buffer is allocated, so that
    the page-locked buffers of the end-to-end path are first touched on the GPU's own NUMA node: with 8 ranks on a
    two-socket box the device->host copies otherwise cross the socket interconnect.  Returns the number of cores bound
    to, or None when the topology is not available (single socket, restricted cpuset, no NVML)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus or cpus == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def emit(line: str):
    """The one JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse_args()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and "RANK" not in os.environ:
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "bench.py")] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args, rank)
        return

    numa_cores = bind_to_gpu_numa_node(local_rank) if world > 1 else None

    import numpy as np
    import torch
    import torch.distributed as dist

    import pkgpath

    pkgpath.load()
    from mapping_private_b200 import cab, shard, synth

    if args.workload == "grsd":
        run_grsd(args, rank, world, local_rank)
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.points
    pts = synth.room(n)
    ctx = cab.Context(local_rank, exact=args.exact)
    ctx.set_shard(rank, world)
    d_xyz = torch.from_numpy(pts).to(dev)  # resident in HBM before the timed region
    torch.cuda.synchronize()
    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    ranges = None

    def exchange(which, width):
        """Concatenate the shards' results: every rank broadcasts its slice in place (NCCL/NVLink)."""
        buf = torch.as_tensor(_DevArray(ctx.device_ptr(which), (n, width)), device=dev)
        shard.exchange_slices(buf, ranges)
        torch.cuda.synchronize()

    stage_s = {"build": 0.0, "normals": 0.0, "unused": 0.0, "rsd": 0.0, "gather": 0.0}

    def step():
        t0 = time.perf_counter()
        ctx.set_cloud_device(d_xyz.data_ptr(), n, 3)
        ctx.build_grid(RADIUS)
        t1 = time.perf_counter()
        ctx.normals(RADIUS, download=False)
        t2 = time.perf_counter()
        t3 = time.perf_counter()  # no exchange between the passes: halo normals are recomputed locally
        ctx.rsd(RADIUS, ndiv=NDIV, plane_radius=PLANE_RADIUS, download=False)
        t4 = time.perf_counter()
        if world > 1 and args.gather:  # optional: every rank ends up with all results in HBM
            exchange(cab.BUF_NRM_SORTED, 4)
            exchange(cab.BUF_RSD_SORTED, 2)
        t5 = time.perf_counter()
        for k, v in zip(stage_s, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            stage_s[k] += v

    # shard ranges are a deterministic function of the cloud; exchange them once
    ctx.set_cloud_device(d_xyz.data_ptr(), n, 3)
    ctx.build_grid(RADIUS)
    if world > 1:
        mine = ctx.shard_range()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        ranges = gathered
    for _ in range(max(args.warmup, 3)):
        step()

    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.profile()["kernel_launches"]
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = {"build_ms": [], "normals_ms": [], "rsd_ms": []}
    for k in stage_s:
        stage_s[k] = 0.0
    ev0.record(lib_stream)
    for _ in range(args.steps):
        step()
        p = ctx.profile()
        for k in phases:
            phases[k].append(p[k])
    ev1.record(lib_stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    prof = ctx.profile()
    launches = prof["kernel_launches"] - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        ks = torch.tensor([prof["neighbour_sum"], launches], device=dev, dtype=torch.int64)
        dist.all_reduce(ks, op=dist.ReduceOp.SUM)
        neighbour_sum, launches = int(ks[0].item()), int(ks[1].item())
    else:
        neighbour_sum = prof["neighbour_sum"]
    ms_per_step = elapsed_ms / args.steps
    value = n / (ms_per_step * 1e-3)
    kbar = neighbour_sum / n
    my_stages = {k: 1e3 * v / args.steps for k, v in stage_s.items()}  # host wall per stage (each stage syncs)
    per_rank = [my_stages]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, my_stages)

    # ---- roofline of the dominant kernel (rsd_kernel), SURVEY section 8(d) accounting -------
    peak, peak_src = hbm_peak()
    rsd_ms = statistics.mean(phases["rsd_ms"])
    nrm_ms = statistics.mean(phases["normals_ms"])
    build_ms = statistics.mean(phases["build_ms"])
    shard_pts = n / world
    shard_k = prof["neighbour_sum"]  # this rank's queries
    rsd_bytes = 32.0 * shard_k + 40.0 * shard_pts
    nrm_bytes = 16.0 * shard_k + 32.0 * shard_pts
    achieved = rsd_bytes / (rsd_ms * 1e-3) / 1e9
    traffic = None
    tpath = ROOT / "profiles" / "traffic.json"
    if tpath.exists():
        try:
            traffic = json.loads(tpath.read_text()).get("rsd_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    step_bytes = n * (48.0 * kbar + 72.0 + 120.0)
    roofline = {"bound": "hbm", "kernel": "rsd_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": rsd_bytes, "kernel_ms": rsd_ms,
                "normals_kernel": {"achieved": nrm_bytes / (nrm_ms * 1e-3) / 1e9, "kernel_ms": nrm_ms,
                                   "frac": nrm_bytes / (nrm_ms * 1e-3) / 1e9 / peak},
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ---- end to end through the C ABI with pinned host buffers ------------------------------
    e2e = None
    if not args.no_e2e:
        h_xyz = torch.from_numpy(pts).pin_memory()
        h_n4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        h_rmin = torch.empty(n if world == 1 else 2 * n, dtype=torch.float32).pin_memory()
        h_rmax = torch.empty(n, dtype=torch.float32).pin_memory()
        h_idx = torch.empty(n, dtype=torch.int32).pin_memory()
        L = cab.lib()
        import ctypes as C

        def fp(t):
            return C.cast(t.data_ptr(), C.POINTER(C.c_float))

        vp0 = (C.c_float * 3)(0.0, 0.0, 0.0)

        my_lo, my_hi = shard.split_range(n, world)[rank]
        d_full = torch.empty((n, 3), dtype=torch.float32, device=dev) if world > 1 else None
        d_slice = torch.empty((my_hi - my_lo, 3), dtype=torch.float32, device=dev) if world > 1 else None

        e2e_stage = {"upload": 0.0, "build": 0.0, "normals_rsd_d2h": 0.0}

        def e2e_step():
            t0 = time.perf_counter()
            if world == 1:
                ctx._check(L.cab_upload_cloud(ctx._h, fp(h_xyz), C.c_int64(n), C.c_int32(3)), "cab_upload_cloud")
                ctx.n = n
            else:
                # every rank uploads 1/world of the cloud over PCIe, the slices are all-gathered over NVLink
                d_slice.copy_(h_xyz[my_lo:my_hi], non_blocking=True)
                shard.gather_cloud(d_full, d_slice, rank, world)
                torch.cuda.synchronize()
                ctx.set_cloud_device(d_full.data_ptr(), n, 3)
            t1 = time.perf_counter()
            ctx.build_grid(RADIUS)
            t2 = time.perf_counter()
            # both passes in one call: the normals leave on the copy stream while the RSD kernel runs.
            # world > 1: each rank returns its own slice (sorted order) plus the input indices it belongs to
            ctx._check(L.cab_normals_rsd(ctx._h, C.c_double(RADIUS), C.c_int32(0), vp0, C.c_int32(0), C.c_int32(NDIV),
                                         C.c_double(PLANE_RADIUS), C.c_int32(0), C.c_int32(0 if world == 1 else 1), fp(h_n4),
                                         fp(h_rmin), fp(h_rmax) if world == 1 else None,
                                         C.cast(h_idx.data_ptr(), C.POINTER(C.c_int32)) if world > 1 else None), "cab_normals_rsd")
            t3 = time.perf_counter()
            for k, v in zip(e2e_stage, (t1 - t0, t2 - t1, t3 - t2)):
                e2e_stage[k] += v

        e2e_steps = max(2, min(args.steps, 5))
        e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for k in e2e_stage:
            e2e_stage[k] = 0.0
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n / (dt / e2e_steps), "unit": UNIT, "h2d_bytes_per_step": int(n * 12),  # summed over ranks
               "d2h_bytes_per_step": int(n * 24) if world == 1 else int(n * 28), "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "stages_ms_rank0": {k: 1e3 * v / e2e_steps for k, v in e2e_stage.items()},
               "path": ("cab_upload_cloud -> cab_build_grid -> cab_normals_rsd (normals D2H overlaps the RSD kernel), pinned host buffers" if world == 1 else
                        "per rank: H2D of 1/N of the cloud + NCCL all-gather of the slices -> cab_set_cloud_device -> cab_build_grid -> "
                        "cab_normals_rsd, CAB_OUT_SHARD_SORTED (own slice + input indices), pinned host buffers; bytes summed over ranks")}

        # Supplementary (N = 1): two frames in flight.  Two contexts (each its own stream and device arena) take
        # alternate frames from two host threads, so the H2D / D2H copies of one frame overlap the kernels of the
        # other -- every copy is still inside the timed region.  `value` above stays the one-frame-at-a-time number.
        if world == 1:
            import threading

            ctx_b = cab.Context(local_rank, exact=args.exact)
            bufs_b = [torch.empty_like(h_n4).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
                      torch.empty(n, dtype=torch.float32).pin_memory()]

            def frame(c, o4, omin, omax):
                c._check(L.cab_upload_cloud(c._h, fp(h_xyz), C.c_int64(n), C.c_int32(3)), "cab_upload_cloud")
                c.n = n
                c.build_grid(RADIUS)
                c._check(L.cab_normals_rsd(c._h, C.c_double(RADIUS), C.c_int32(0), vp0, C.c_int32(0), C.c_int32(NDIV),
                                           C.c_double(PLANE_RADIUS), C.c_int32(0), C.c_int32(0), fp(o4), fp(omin), fp(omax), None),
                         "cab_normals_rsd")

            lanes = [(ctx, h_n4, h_rmin, h_rmax), (ctx_b, *bufs_b)]
            frames_each = max(2, e2e_steps // 2 + 1)
            for lane in lanes:
                frame(*lane)  # warm-up (allocates the second context's arena)
            torch.cuda.synchronize()
            gate = threading.Barrier(3)

            def worker(lane):
                gate.wait()
                for _ in range(frames_each):
                    frame(*lane)

            threads = [threading.Thread(target=worker, args=(lane,)) for lane in lanes]
            for t in threads:
                t.start()
            gate.wait()
            t0 = time.perf_counter()
            for t in threads:
                t.join()
            torch.cuda.synchronize()
            dt2 = time.perf_counter() - t0
            same = bool(torch.equal(h_rmin, bufs_b[1]) and torch.equal(h_n4.view(torch.int32), bufs_b[0].view(torch.int32)))
            e2e["two_frames_in_flight"] = {"value": 2 * frames_each * n / dt2, "unit": UNIT, "ms_per_frame": 1e3 * dt2 / (2 * frames_each),
                                           "frames": 2 * frames_each, "results_identical": same,
                                           "path": "two contexts / streams, alternate frames from two host threads: copies of one frame overlap the kernels of the other"}
            ctx_b.close()

    # ---- CPU baseline (oracle port) on rank 0 at N = 1 ---------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = slab_sample(pts, args.cpu_sample)
        rate, dt, cores = cpu_points_per_s(sample)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
               "sample": f"x-slab of the room cloud, {sample.shape[0]} points (same density), oracle normals+RSD streaming mode"}
        try:
            parity = parity_gate(ctx, sample)
        except Exception as e:  # a failed gate must show up in the line, not kill the measurement
            parity = {"error": repr(e)}
        try:
            cpu["c1_reference_faithful"] = c1_reference_faithful(ctx)
        except Exception as e:
            cpu["c1_reference_faithful"] = {"error": repr(e)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if not args.exact else "f64", "data": "synthetic",
            "config": {"workload": "C4 20M-point synthetic room, normals+RSD r=2cm, query-sharded", "points": n,
                       "radius_m": RADIUS, "distance_div": NDIV, "plane_radius": PLANE_RADIUS, "max_nn": "unlimited",
                       "mean_neighbours": kbar, "candidates_tested_per_query": prof["candidate_sum"] / max(1.0, n / world), "l2": "inputs_larger_than_l2 (pos+normals 640 MB vs 126 MB L2)",
                       "parallelism": (f"query-shard x{world}, cloud+grid replicated, halo normals recomputed locally, results stay sharded" + (" then all-gathered (NCCL)" if args.gather else "")) if world > 1 else "single GPU",
                       "mode": "exact-fp64" if args.exact else "fast-fp32",
                       "host_affinity": (f"rank 0 bound to the {numa_cores} cores of its GPU's NUMA node" if numa_cores else "not bound")},
            "phases_ms": {"build": build_ms, "normals": nrm_ms, "rsd": rsd_ms},
            "per_rank_stage_ms": per_rank,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
