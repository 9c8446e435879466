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
KERNELS_VERSION = "r02-v3"  # bumped with every kernel change: profiles/traffic.json is only quoted for the kernels it measured


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
    ap.add_argument("--no-extras", action="store_true", help="N=1: skip the other configurations / modes (C2, C3, exact, max_nn=150)")
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
    thread every 10 ms, so that even a 60 ms region (8 GPUs) gets samples; nvidia-smi -lms is the fallback.  Only rank 0
    polls (its line is the one printed), the constant maximum clock is read once, and the first query of every kind is
    made before the region starts: NVML queries go through the driver's node-wide lock, and eight ranks polling four
    queries every 4 ms were a measurable source of launch jitter in a 3 ms step."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_s=0.010, enabled=True):
        self.index = index
        self.period_s = period_s
        self.enabled = enabled
        self.proc = None
        self.thread = None
        self.rows = []  # (sm MHz, max MHz, watts, reasons bit mask)
        self._stop = False
        self.nvml = None
        self.max_mhz = None
        if not enabled:
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._sample()  # the first query of every kind outside the timed region
            self.rows.clear()
        except Exception:
            self.nvml = None

    def _sample(self, with_power=False):
        nv = self.nvml
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
            mx = self.max_mhz
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            watts = None
            if with_power:  # (the power query is the slow one: asked once when the region is over, not while it runs)
                try:
                    watts = nv.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                except Exception:
                    watts = None
            self.rows.append((float(sm), float(mx) if mx is not None else float(sm), watts, int(mask)))
        except Exception:
            pass

    def _poll(self):
        while not self._stop:
            self._sample()
            time.sleep(self.period_s)

    def start(self):
        if not self.enabled:
            return
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
        if not self.enabled:
            return None
        if self.nvml:
            self._stop = True
            if self.thread:
                self.thread.join(timeout=2)
            self._sample(with_power=True)  # one last sample, with the board power, right behind the region
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


def host_threads():
    """Threads for the CPU arm: every core this process may run on, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its workers; the reference arm must not inherit that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_points_per_s(pts, nthreads=0, radius=RADIUS):
    """Oracle (CPU restatement) normals + RSD on the given sample, all host threads."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import pyoracle

    pyoracle.build()
    nthreads = nthreads if nthreads > 0 else host_threads()
    t0 = time.perf_counter()
    n4, gap = pyoracle.normals_gap(pts, radius, nthreads=nthreads)
    rmin, rmax, _ = pyoracle.rsd(pts, n4, radius, ndiv=NDIV, plane_radius=PLANE_RADIUS, nthreads=nthreads)
    dt = time.perf_counter() - t0
    cpu_points_per_s.last = (n4, gap, rmin, rmax)  # kept for the parity gate of the same run
    return pts.shape[0] / dt, dt, nthreads


def c1_reference_faithful(ctx):
    """BASELINE config C1 (the reference's own CPU-runnable case: sample_pipeline's RSD on a 100 k-point tabletop cloud,
    r = 2 cm, the plugin's default max_nn = 150) the way the reference runs it -- kd-tree, materialised neighbour lists,
    estimation loop, ONE thread, the three phases timed where radius_estimation.cpp:108,125,216 log them -- next to the
    same call through the C ABI (upload + grid + thresholds + RSD + download, host buffers).  A reported baseline."""
    import numpy as np

    import pyoracle
    from mapping_private_b200 import synth

    pts = synth.tabletop(100_000)
    n4, _ = pyoracle.normals(pts, RADIUS, nthreads=host_threads())
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


def neighbour_sets_equal(ctx, pts, radius, n_queries=65536):
    """neighbour index SETS, bit-exact, on a 64 k-query sample (a contiguous range of the input order, which
    is a random spatial sample of the cloud): cab_neighbors_debug against the oracle's radius search."""
    import numpy as np

    import pyoracle

    nq = min(n_queries, pts.shape[0])
    q0 = (pts.shape[0] - nq) // 2
    goff, gidx, gd2 = ctx.neighbors(radius, q0, q0 + nq)
    ooff, oidx, od2 = pyoracle.radius_search(pts, pts[q0:q0 + nq], radius, nthreads=host_threads())
    same_counts = bool(np.array_equal(goff, ooff))
    same_sets = same_d2 = False
    if same_counts:
        qid = np.repeat(np.arange(nq), np.diff(goff))
        go = np.lexsort((gidx, qid))
        oo = np.lexsort((oidx, qid))
        same_sets = bool(np.array_equal(gidx[go], oidx[oo]))
        same_d2 = bool(np.array_equal(gd2[go].view(np.uint32), od2[oo].view(np.uint32)))
    return {"queries": int(nq), "pairs": int(ooff[-1]), "per_query_counts_equal": same_counts, "index_sets_equal": same_sets,
            "d2_bits_equal": same_d2}


def radii_rel(gmin, gmax, omin, omax):
    import numpy as np

    return np.maximum(np.abs(gmin - omin) / omin, np.abs(gmax - omax) / omax)


def parity_gate(ctx_fast, ctx_exact, sample, radius=RADIUS, sens_points=1_000_000):
    """the parity gates that go with every benchmark line, in the mode that is benchmarked AND in exact mode.
    The CUDA path on the slab the CPU baseline just processed, against that run's oracle results:
      * neighbour sets bit-exact on a 64 k-query sample, per-query counts over the whole slab through the sum;
      * normals within 1e-4 rad (sign-insensitive), with the oracle's own conditioning (eigenvalue gap) of the worst points;
      * radii END TO END (device normals -> device RSD against oracle normals -> oracle RSD) within 1e-4 relative, the
        tail listed instead of dropped, next to what the reference's own arithmetic does to a 1-ulp change of the normals
        (its cosine is an fp32 expression, radius_estimation.cpp:153-155: near |cos| = 1 one ulp of a normal moves the
        angle by up to 3.5e-4 rad), and the same radii given identical normals.
    The oracle is the checker here, nothing of it is timed."""
    import numpy as np

    import pyoracle

    o4, gap, omin, omax = cpu_points_per_s.last
    out = {"sample_points": int(sample.shape[0]), "tolerance": {"normals_rad": 1e-4, "radii_rel": 1e-4}}
    good_o = ~np.isnan(o4[:, 0])

    def one_mode(c, name):
        c.upload(sample)
        c.build_grid(radius)
        g4 = c.normals(radius)
        k_sum = int(c.profile()["neighbour_sum"])
        gmin, gmax = c.rsd(radius, ndiv=NDIV, plane_radius=PLANE_RADIUS)  # end to end: the device's own normals
        good = good_o & ~np.isnan(g4[:, 0])
        sin_angle = np.zeros(sample.shape[0])
        sin_angle[good] = np.linalg.norm(np.cross(g4[good, :3].astype(np.float64), o4[good, :3].astype(np.float64)), axis=1)
        rel = radii_rel(gmin, gmax, omin, omax)
        over = rel > 1e-4
        worst = int(np.argmax(sin_angle))
        c.set_normals(o4)
        smin, smax = c.rsd(radius, ndiv=NDIV, plane_radius=PLANE_RADIUS)
        res = {"nan_normals_equal": bool(np.array_equal(np.isnan(o4[:, 0]), np.isnan(g4[:, 0]))),
               "normals_fraction_over_1e-4_rad": float(np.mean(sin_angle > 1e-4)), "normals_max_rad": float(sin_angle.max()),
               "normals_p99.9_rad": float(np.percentile(sin_angle, 99.9)),
               "normals_worst_point_eigen_gap": float(gap[worst]),
               "normals_over_1e-4_with_gap_below_0.01": int(np.sum((sin_angle > 1e-4) & (gap < 0.01))),
               "normals_over_1e-4_total": int(np.sum(sin_angle > 1e-4)),
               "radii_e2e_fraction_over_1e-4": float(np.mean(over)), "radii_e2e_over_1e-4": int(over.sum()),
               "radii_e2e_max_rel": float(np.nanmax(rel)), "radii_e2e_p99.9_rel": float(np.percentile(rel, 99.9)),
               "radii_max_rel_given_same_normals": float(np.nanmax(radii_rel(smin, smax, omin, omax)))}
        return res, k_sum, over

    fast, k_fast, over_fast = one_mode(ctx_fast, "fast")
    out["neighbour_sets_64k"] = neighbour_sets_equal(ctx_fast, sample, radius)
    out["fast-fp32"] = fast
    if ctx_exact is not None:
        out["exact-fp64"], k_exact, _ = one_mode(ctx_exact, "exact")
        out["neighbour_sum_fast_equals_exact"] = bool(k_fast == k_exact)
    # what the reference's own arithmetic does with normals that differ in the last bit: the oracle against itself
    m = min(sens_points, sample.shape[0])
    sub = np.ascontiguousarray(sample[:m]) if m == sample.shape[0] else slab_sample(sample, m)
    s4, _ = pyoracle.normals(sub, radius, nthreads=host_threads())
    rng = np.random.default_rng(0xC10D)
    bumped = s4.copy()
    sign = rng.integers(0, 2, size=(sub.shape[0], 3)).astype(np.float32) * 2 - 1
    bumped[:, :3] = np.nextafter(s4[:, :3], s4[:, :3] + sign)
    amin, amax, _ = pyoracle.rsd(sub, s4, radius, ndiv=NDIV, plane_radius=PLANE_RADIUS, nthreads=host_threads())
    bmin, bmax, _ = pyoracle.rsd(sub, bumped, radius, ndiv=NDIV, plane_radius=PLANE_RADIUS, nthreads=host_threads())
    self_rel = radii_rel(bmin, bmax, amin, amax)
    out["oracle_vs_oracle_with_normals_1ulp_off"] = {
        "points": int(sub.shape[0]), "radii_fraction_over_1e-4": float(np.mean(self_rel > 1e-4)),
        "radii_max_rel": float(np.nanmax(self_rel)), "radii_p99.9_rel": float(np.percentile(self_rel, 99.9)),
        "meaning": "the reference's fp32 cosine (radius_estimation.cpp:153-155) turns a 1-ulp change of the normals into this "
                   "spread of the radii: the conditioning margin of the end-to-end radii tolerance"}
    return out


def run_reference(args, rank):
    """--impl reference: the reference's algorithm on the host cores (oracle port: the reference
    plugins need ROS/PCL/ANN, none of which exist here -- DESIGN.md "Oracle").  Under torchrun rank 0 alone runs it, with
    every core of the box (torchrun's OMP_NUM_THREADS=1 is not inherited: the thread count is set explicitly)."""
    if rank != 0:
        return
    import pkgpath

    pkgpath.load()
    from mapping_private_b200 import synth

    pts = synth.room(args.points)  # full cloud so that the slab has the workload's density
    sample = slab_sample(pts, 1_500_000)  # ~2-4 s of work per step on 16 cores: K + W steps stay within minutes
    threads = host_threads()
    rates = []
    t_start = time.perf_counter()
    steps_done = 0
    for s in range(args.warmup + args.steps):
        rate, dt, cores = cpu_points_per_s(sample, nthreads=threads)
        if s >= args.warmup:
            rates.append((rate, dt))
            steps_done += 1
        if time.perf_counter() - t_start > 420 and steps_done >= 2:  # a slow box must not run into the driver's limit
            break
    value = len(rates) * sample.shape[0] / sum(dt for _, dt in rates)
    ms = 1e3 * sum(dt for _, dt in rates) / len(rates)
    desc = f"x-slab of the room cloud, {sample.shape[0]} points per step (same density as the {args.points}-point workload)"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 20M-point synthetic room, normals+RSD r=2cm (bounded sample per step)", "points": args.points,
                   "radius_m": RADIUS, "distance_div": NDIV, "plane_radius": PLANE_RADIUS, "max_nn": "unlimited"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def grsd_c3(ctx, n_clusters, reps, cpu_check=48):
    """Config C3 on one GPU: GRSD-21 of a batch of segmented clusters (2.5 cm voxels), host buffers in, 21 int32 bins per
    cluster out (the GlobalRSD plugin's work).  Returns the numbers that go into the bench line."""
    import numpy as np
    import torch

    from mapping_private_b200 import synth

    xyz, off = synth.clusters(n_clusters)
    xyz = torch.from_numpy(np.ascontiguousarray(xyz)).pin_memory().numpy()
    leaf = 0.025
    hist = ctx.grsd_batch(xyz, off, leaf, r_normals=0.02)  # warm-up
    l0 = ctx.profile()["kernel_launches"]
    t0 = time.perf_counter()
    kern = []
    for _ in range(reps):
        hist = ctx.grsd_batch(xyz, off, leaf, r_normals=0.02)
        p = ctx.profile()
        kern.append(p["build_ms"] + p["normals_ms"] + p["grsd_ms"])
    dt = (time.perf_counter() - t0) / reps
    launches = (ctx.profile()["kernel_launches"] - l0) // reps
    vox = ctx.grsd_voxels(n_clusters)
    n_pts, n_vox = int(off[-1]), int(vox["offsets"][-1])
    k_n = ctx.profile()["neighbour_sum"] / max(1, n_pts)  # normals pass of the batch
    # SURVEY 8(d), GRSD per cluster: Nc(16 k_n + 32) + Nc(16 + 32) + V(32 k_r + 40) + V 26 4 + 84, with k_r ~ the points of
    # a 2.5 cm voxel's RSD sphere, bounded by the cluster's points per voxel times the sphere / voxel volume ratio
    k_r = (n_pts / max(1, n_vox)) * (4.0 / 3.0 * 3.14159265 * 0.0216506 ** 3) / leaf ** 3 * 2.0
    alg = n_pts * (16.0 * k_n + 32) + n_pts * 48.0 + n_vox * (32.0 * k_r + 40) + n_vox * 104.0 + 84.0 * n_clusters
    peak, _ = hbm_peak()
    out = {"workload": f"C3 GRSD-21 on {n_clusters} synthetic clusters, leaf 2.5 cm, normals r=2cm, exact mode, pinned host buffers in, histograms out",
           "clouds_per_s": n_clusters / dt, "ms_per_batch": 1e3 * dt, "kernels_ms_per_batch": statistics.mean(kern), "points": n_pts,
           "voxels": n_vox, "gpu_launches_per_batch": int(launches),
           "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / (statistics.mean(kern) * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": alg / (statistics.mean(kern) * 1e-3) / 1e9 / peak,
                        "note": "SURVEY 8(d) GRSD formula; the batch is latency / fp64 bound, not HBM bound (profiles/)"}}
    if cpu_check:
        sys.path.insert(0, str(ROOT / "oracle"))
        import pyoracle

        pyoracle.build()
        ncheck = min(n_clusters, cpu_check)
        t1 = time.perf_counter()
        bad = 0
        for c in range(ncheck):
            o = pyoracle.grsd21(xyz[off[c]:off[c + 1]], leaf, r_normals=0.02, nthreads=host_threads())
            bad += int(not np.array_equal(o["hist21"], hist[c]))
        cdt = time.perf_counter() - t1
        out["cpu_baseline"] = {"value": ncheck / cdt, "unit": "clouds/s", "cores": host_threads(), "kind": "port",
                               "sample": f"first {ncheck} clusters, oracle normals + voxel RSD + GRSD"}
        out["histograms_differing_from_oracle"] = int(bad)
        out["histograms_checked"] = int(ncheck)
    return out, hist


def run_grsd(args, rank, world, local_rank):
    """Config C3: GRSD-21 of a batch of segmented clusters (2.5 cm voxels), cluster-per-GPU.
    Host buffers in, 21 int32 bins per cluster out (the GlobalRSD plugin's work); the histograms of the
    ranks are summed with one small integer all-reduce behind the C ABI (cab_comm_allreduce_i32, NCCL)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from mapping_private_b200 import cab, shard, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = cab.Context(local_rank, exact=True)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ctx.comm_init(broadcast_comm_id(rank, cab), rank, world)
    xyz, off = synth.clusters(args.clusters)
    sizes = np.diff(off)
    mine = shard.assign_clusters_lpt(sizes.tolist(), world)[rank]
    my_xyz = np.concatenate([xyz[off[c]:off[c + 1]] for c in mine]) if mine else np.zeros((0, 3), np.float32)
    my_xyz = torch.from_numpy(np.ascontiguousarray(my_xyz)).pin_memory().numpy()  # page-locked host buffer (H2D at PCIe speed)
    my_off = np.concatenate([[0], np.cumsum(sizes[mine])]).astype(np.int32)
    leaf = 0.025

    def step():
        hist = np.zeros((args.clusters, 21), np.int32)
        if mine:
            hist[mine] = ctx.grsd_batch(my_xyz, my_off, leaf, r_normals=0.02)
        if world > 1:
            ctx.comm_allreduce_i32(hist)
        return hist

    for _ in range(max(args.warmup, 3)):
        hist = step()
    sampler = ClockSampler(local_rank, enabled=(rank == 0))
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
            bad += int(not np.array_equal(o["hist21"], hist[c]))
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
                       "points": int(off[-1]), "leaf_m": leaf,
                       "parallelism": f"clusters LPT x{world}, one int32 all-reduce of the histograms behind the C ABI (cab_comm_allreduce_i32, NCCL)",
                       "timed": "pinned host buffers in, histograms out (H2D + D2H inside)"},
            "kernels_ms_per_step_rank0": statistics.mean(kern_ms), "cpu_baseline": cpu, "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": args.clusters / (dt / args.steps), "unit": "clouds/s", "h2d_bytes_per_step": int(off[-1]) * 12,
                    "d2h_bytes_per_step": args.clusters * 84},
        }))
    if world > 1:
        dist.destroy_process_group()


def broadcast_comm_id(rank, cab):
    """cab_comm_get_id on rank 0 (ncclGetUniqueId behind the C ABI); torch.distributed only carries the 128 bytes."""
    import torch.distributed as dist

    box = [cab.comm_get_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


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
    if args.gpus > 1 and "RANK" not in os.environ:  # before stdout is redirected: the workers inherit the real one
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "bench.py")] + sys.argv[1:])
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    numa_cores = bind_to_gpu_numa_node(local_rank) if world > 1 else None

    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import pkgpath

    pkgpath.load()
    from mapping_private_b200 import cab, synth

    if args.workload == "grsd":
        run_grsd(args, rank, world, local_rank)
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = cab.Context(local_rank, exact=args.exact)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # the group behind the C ABI: NCCL carries the bootstrap, the data plane is peer memory (csrc/cab_comm.cu)
        ctx.comm_init(broadcast_comm_id(rank, cab), rank, world)
        # every result crosses NVLink once, to the rank that owns the point's input index: after a step rank g holds the
        # channels of input indices [n g / N, n (g + 1) / N) in input order
        ctx.comm_set_layout(cab.COMM_LAYOUT_INPUT_RANGES)
    n = args.points
    pts = synth.room(n)
    d_xyz = torch.from_numpy(pts).to(dev)  # resident in HBM before the timed region
    torch.cuda.synchronize()
    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    warmup = max(args.warmup, 3)

    def step(c=ctx, max_nn_rsd=0):
        """One pass of the hot path over the cloud in ONE C call: grid build + normals + RSD and, in a group, the
        concatenation of the ranks' results (input order, rank g holding the g-th N-th of the input indices)."""
        c.set_cloud_device(d_xyz.data_ptr(), n, 3)
        # at every N the step leaves the channels of radius_estimation.cpp:204-214 in INPUT order on the device: a group's
        # ranks hold one input range each, a single context all of it (CAB_STEP_INPUT_ORDER: scattered by the RSD kernel)
        c.step_normals_rsd(RADIUS, RADIUS, max_nn_rsd=max_nn_rsd, ndiv=NDIV, plane_radius=PLANE_RADIUS,
                           flags=cab.STEP_INPUT_ORDER if world == 1 else 0)

    for _ in range(warmup):
        step()

    sampler = ClockSampler(local_rank, enabled=(rank == 0))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.profile()["kernel_launches"]
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = {"build_ms": [], "normals_ms": [], "rsd_ms": [], "exchange_ms": [], "step_ms": []}
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
    my_phase = {k: statistics.mean(v) for k, v in phases.items()}
    # (a single slow step -- a host hiccup on one rank -- shows here instead of hiding in the mean)
    my_phase["step_ms_median"] = statistics.median(phases["step_ms"])
    my_phase["step_ms_max"] = max(phases["step_ms"])
    my_phase["build_ms_max"] = max(phases["build_ms"])
    my_phase["own_queries"] = (lambda b_e: b_e[1] - b_e[0])(ctx.shard_range())
    my_phase["points_sorted"] = prof["n_sorted"]
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        ks = torch.tensor([prof["neighbour_sum"], launches, prof["candidate_sum"]], device=dev, dtype=torch.int64)
        dist.all_reduce(ks, op=dist.ReduceOp.SUM)
        neighbour_sum, launches, candidate_sum = int(ks[0].item()), int(ks[1].item()), int(ks[2].item())
        per_rank = [None] * world
        dist.all_gather_object(per_rank, my_phase)
    else:
        neighbour_sum, candidate_sum = prof["neighbour_sum"], prof["candidate_sum"]
        per_rank = [my_phase]
    ms_per_step = elapsed_ms / args.steps
    value = n / (ms_per_step * 1e-3)
    kbar = neighbour_sum / n

    concatenated = None
    if world == 1:
        # the input-order arrays the timed step left on the device against the sorted arrays + permutation (cab_download)
        p4 = ctx.device_ptr(cab.BUF_NRM_INPUT_RANGE)
        p2 = ctx.device_ptr(cab.BUF_RSD_INPUT_RANGE)
        in4 = torch.as_tensor(_DevArray(p4, (n, 4), "<i4"), device=dev)
        in2 = torch.as_tensor(_DevArray(p2, (n, 2), "<i4"), device=dev)
        f4, fmin, fmax = ctx.download()
        same = bool(np.array_equal(in4.cpu().numpy(), f4.view(np.int32)) and np.array_equal(in2[:, 0].cpu().numpy(), fmin.view(np.int32))
                    and np.array_equal(in2[:, 1].cpu().numpy(), fmax.view(np.int32)))
        concatenated = {"layout": "input order on the device (one context holds all points)", "equals_cab_download": same, "points": n}
        del in4, in2, f4, fmin, fmax
    if world > 1:
        # the concatenation is part of the timed step: poison every rank's range, run one more step, and count the
        # entries the ranks' RSD kernels (and, for non-finite points, the owner's selection pass) have written
        p4, cnt = ctx.comm_device_ptr(cab.BUF_NRM_INPUT_RANGE)
        p2, _ = ctx.comm_device_ptr(cab.BUF_RSD_INPUT_RANGE)
        own4 = torch.as_tensor(_DevArray(p4, (cnt * 4,), "<i4"), device=dev)
        own2 = torch.as_tensor(_DevArray(p2, (cnt * 2,), "<i4"), device=dev)
        own4.fill_(-1)
        own2.fill_(-1)
        torch.cuda.synchronize()
        dist.barrier()
        step()
        torch.cuda.synchronize()
        dist.barrier()
        written = torch.tensor([int(((own4.view(-1, 4) != -1).any(dim=1) & (own2.view(-1, 2) != -1).any(dim=1)).sum().item()), cnt],
                               device=dev, dtype=torch.int64)
        dist.all_reduce(written, op=dist.ReduceOp.SUM)
        concatenated = {"layout": "input ranges: rank g holds input indices [n g / N, n (g + 1) / N) in input order",
                        "entries_written_by_one_step": int(written[0].item()), "entries": int(written[1].item()), "points": n}

    # ---- roofline of the dominant kernel (rsd_kernel), SURVEY section 8(d) accounting -------
    peak, peak_src = hbm_peak()
    rsd_ms, nrm_ms, build_ms = my_phase["rsd_ms"], my_phase["normals_ms"], my_phase["build_ms"]
    shard_pts = my_phase["own_queries"]
    shard_k = prof["neighbour_sum"]  # this rank's queries
    rsd_bytes = 32.0 * shard_k + 40.0 * shard_pts
    nrm_bytes = 16.0 * shard_k + 32.0 * shard_pts
    achieved = rsd_bytes / (rsd_ms * 1e-3) / 1e9
    traffic, traffic_note = None, "not measured for this configuration"
    tpath = ROOT / "profiles" / "traffic.json"
    if tpath.exists() and world == 1 and not args.exact:
        try:
            tj = json.loads(tpath.read_text())
            if int(tj.get("points", 0)) == n and tj.get("kernels") == KERNELS_VERSION:
                traffic = tj.get("rsd_kernel_dram_bytes_per_launch")
                traffic_note = tj.get("source", "")
        except Exception:
            traffic = None
    step_bytes = n * (48.0 * kbar + 72.0 + 120.0)
    roofline = {"bound": "hbm", "kernel": "rsd_fast_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": rsd_bytes, "kernel_ms": rsd_ms,
                "binding_resource": "instruction issue and the shared-memory pipe, not DRAM: the x-sorted rows turn the neighbour "
                                    "gather into shared-memory / L2 hits (ncu in profiles/: dram throughput < 1 % of peak, issue active "
                                    "and LSU wavefronts 80-90 %); the HBM figure is the SURVEY 8(d) byte model over the kernel time",
                "normals_kernel": {"achieved": nrm_bytes / (nrm_ms * 1e-3) / 1e9, "kernel_ms": nrm_ms,
                                   "frac": nrm_bytes / (nrm_ms * 1e-3) / 1e9 / peak},
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ---- end to end through the C ABI with pinned host buffers ------------------------------
    e2e = None
    if not args.no_e2e:
        L = cab.lib()
        h_xyz = torch.from_numpy(pts).pin_memory()
        vp0 = (C.c_float * 3)(0.0, 0.0, 0.0)

        def fp(t, off_elems=0):
            return C.cast(t.data_ptr() + 4 * off_elems, C.POINTER(C.c_float))

        if world == 1:
            h_n4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
            h_rmin = torch.empty(n, dtype=torch.float32).pin_memory()
            h_rmax = torch.empty(n, dtype=torch.float32).pin_memory()
            e2e_stage = {"upload": 0.0, "build": 0.0, "normals_rsd_d2h": 0.0}

            def e2e_step():
                t0 = time.perf_counter()
                ctx._check(L.cab_upload_cloud(ctx._h, fp(h_xyz), C.c_int64(n), C.c_int32(3)), "cab_upload_cloud")
                ctx.n = n
                t1 = time.perf_counter()
                ctx.build_grid(RADIUS)
                t2 = time.perf_counter()
                # both passes in one call: the normals leave on the copy stream while the RSD kernel runs
                ctx._check(L.cab_normals_rsd(ctx._h, C.c_double(RADIUS), C.c_int32(0), vp0, C.c_int32(0), C.c_int32(NDIV),
                                             C.c_double(PLANE_RADIUS), C.c_int32(0), C.c_int32(0), fp(h_n4), fp(h_rmin), fp(h_rmax), None),
                           "cab_normals_rsd")
                t3 = time.perf_counter()
                for k, v in zip(e2e_stage, (t1 - t0, t2 - t1, t3 - t2)):
                    e2e_stage[k] += v

            path = "cab_upload_cloud -> cab_build_grid -> cab_normals_rsd (normals D2H overlaps the RSD kernel), pinned host buffers"
            d2h_bytes = n * 24
        else:
            # ONE host array shared by the ranks (POSIX shared memory, page-locked in every process): rank g's device->host
            # copy fills rows [n g / N, n (g + 1) / N) of the channels in INPUT order, as radius_estimation.cpp:204-214 leaves them
            shm_path = f"/dev/shm/cab_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
            total_floats = n * 6
            if rank == 0:
                with open(shm_path, "wb") as f:
                    f.truncate(total_floats * 4)
            dist.barrier()
            shared = np.memmap(shm_path, dtype=np.float32, mode="r+", shape=(total_floats,))
            h_all = torch.from_numpy(shared)
            rc = torch.cuda.cudart().cudaHostRegister(h_all.data_ptr(), total_floats * 4, 0)
            registered = int(rc) == 0 if not isinstance(rc, tuple) else int(rc[0]) == 0
            lo, hi = n * rank // world, n * (rank + 1) // world
            e2e_stage = {"upload_replicate": 0.0, "step": 0.0, "download_range": 0.0}

            def e2e_step():
                t0 = time.perf_counter()
                # every rank uploads 1/N of the cloud over PCIe and copies it to the peers over NVLink (copy engines)
                ctx._check(L.cab_comm_upload_cloud(ctx._h, fp(h_xyz), C.c_int64(n), C.c_int32(3)), "cab_comm_upload_cloud")
                ctx.n = n
                t1 = time.perf_counter()
                ctx.step_normals_rsd(RADIUS, RADIUS, ndiv=NDIV, plane_radius=PLANE_RADIUS)
                t2 = time.perf_counter()
                ctx._check(L.cab_comm_download_range(ctx._h, C.c_int64(lo), C.c_int64(hi), fp(h_all, 4 * lo), fp(h_all, 4 * n + lo),
                                                     fp(h_all, 5 * n + lo)), "cab_comm_download_range")
                t3 = time.perf_counter()
                for k, v in zip(e2e_stage, (t1 - t0, t2 - t1, t3 - t2)):
                    e2e_stage[k] += v

            path = ("per rank: cab_comm_upload_cloud (H2D of 1/N of the cloud + copy-engine replication over NVLink) -> cab_step_normals_rsd "
                    "(slab build + normals + RSD, every result pushed by the RSD kernel to the rank that owns its input index) -> "
                    "cab_comm_download_range: each rank copies ITS input-order range into ONE shared page-locked host array; bytes summed over ranks")
            d2h_bytes = n * 24

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
            dist.barrier()  # the shared array is complete only when every rank has copied its range
            dt_all = time.perf_counter() - t0
            t = torch.tensor([max(dt, dt_all)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n / (dt / e2e_steps), "unit": UNIT, "h2d_bytes_per_step": int(n * 12),  # summed over ranks
               "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "stages_ms_rank0": {k: 1e3 * v / e2e_steps for k, v in e2e_stage.items()}, "path": path}
        if world > 1:
            # the shared array against the OTHER layout: one step that leaves every rank's results on every rank
            ctx.comm_set_layout(cab.COMM_LAYOUT_REPLICATED)
            dist.barrier()
            step()
            dist.barrier()
            if rank == 0:
                f4, fmin, fmax = ctx.comm_download_range(0, n)
                sh = np.asarray(shared)
                e2e["shared_host_array"] = {
                    "page_locked": registered,
                    "equals_replicated_layout_on_rank0": bool(np.array_equal(sh[:4 * n].view(np.uint32), f4.reshape(-1).view(np.uint32)) and
                                                              np.array_equal(sh[4 * n:5 * n].view(np.uint32), fmin.view(np.uint32)) and
                                                              np.array_equal(sh[5 * n:].view(np.uint32), fmax.view(np.uint32)))}
            dist.barrier()
            ctx.comm_set_layout(cab.COMM_LAYOUT_INPUT_RANGES)
            torch.cuda.cudart().cudaHostUnregister(h_all.data_ptr())
            del h_all, shared
            if rank == 0:
                try:
                    os.unlink(shm_path)
                except OSError:
                    pass

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
            del bufs_b, h_n4, h_rmin, h_rmax
        del h_xyz

    # ---- the other BASELINE configurations and modes, N = 1 (extra keys of the same line) -----------------------
    extras = {}
    cpu = None
    parity = None
    # GRSD of the ONE large cloud across the ranks (last stage): slab normals, own voxels' labels, labels merged,
    # one int32 all-reduce of the transition counts -- all behind the C ABI (cab_grsd_cloud); every N
    if not args.no_extras or world > 1:
        try:
            ctx.set_cloud_device(d_xyz.data_ptr(), n, 3)
            h = ctx.grsd_cloud(0.025)  # warm-up (allocations)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                h = ctx.grsd_cloud(0.025)
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
                hh = torch.from_numpy(h.astype(np.int64)).to(dev)
                lo_, hi_ = hh.clone(), hh.clone()
                dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
                same_everywhere = bool(torch.equal(lo_, hi_))
            else:
                same_everywhere = True
            g = {"workload": "GRSD-21 of the whole C4 cloud as one cluster, leaf 2.5 cm, normals r = 2 cm, device-resident cloud",
                 "ms_per_cloud": 1e3 * dt / reps, "voxels": int(ctx.grsd_voxels(1)["labels"].shape[0]),
                 "hist21": [int(x) for x in h], "same_histogram_on_every_rank": same_everywhere,
                 "path": "cab_grsd_cloud: slab grid + normals per rank, own voxels' radii and labels, label merge and int32 all-reduce "
                         "of the transition counts through cab_comm_allreduce_i32 (NCCL)" if world > 1 else "cab_grsd_cloud, one GPU"}
            if world > 1 and rank == 0:  # the unsharded result on this rank's GPU, same mode
                c1 = cab.Context(local_rank, exact=args.exact)
                c1.set_cloud_device(d_xyz.data_ptr(), n, 3)
                c1.n = n
                t0 = time.perf_counter()
                h1 = c1.grsd_cloud(0.025)
                g["single_gpu_ms_first_call"] = 1e3 * (time.perf_counter() - t0)
                g["equals_single_gpu"] = bool(np.array_equal(h1, h))
                c1.close()
            extras["grsd_one_cloud"] = g
        except Exception as e:
            extras["grsd_one_cloud"] = {"error": repr(e)}
        ctx.set_cloud_device(d_xyz.data_ptr(), n, 3)
    if rank == 0 and world == 1 and not args.no_extras:
        def timed_steps(fn, reps):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return 1e3 * (time.perf_counter() - t0) / reps

        ctx_x = cab.Context(local_rank, exact=not args.exact)  # the other arithmetic mode
        other = "exact-fp64" if not args.exact else "fast-fp32"
        try:
            ms_other = timed_steps(lambda: step(ctx_x), 3)
            px = ctx_x.profile()
            extras["other_mode"] = {"mode": other, "ms_per_step": ms_other, "points_per_s": n / (ms_other * 1e-3),
                                    "phases_ms": {"build": px["build_ms"], "normals": px["normals_ms"], "rsd": px["rsd_ms"]},
                                    "neighbour_sum_equal": bool(px["neighbour_sum"] == prof["neighbour_sum"])}
        except Exception as e:
            extras["other_mode"] = {"error": repr(e)}
        try:  # the reference plugin's default truncation (radius_estimation.h:82 max_nn_ = 150)
            ms_150 = timed_steps(lambda: step(ctx, max_nn_rsd=150), 3)
            p150 = ctx.profile()
            extras["max_nn_150"] = {"ms_per_step": ms_150, "points_per_s": n / (ms_150 * 1e-3), "rsd_kernel_ms": p150["rsd_ms"],
                                    "mean_neighbours_rsd": p150["neighbour_sum"] / n,
                                    "config": "normals unlimited (NormalEstimation has no max_nn), RSD max_nn=150 (LocalRadiusEstimation default)"}
        except Exception as e:
            extras["max_nn_150"] = {"error": repr(e)}
        del d_xyz
        torch.cuda.empty_cache()
        try:  # C2: 1 M-point scan, r = 3 cm (the plugin's default radius), range-dependent density
            scan = synth.scan(1_000_000)
            d_scan = torch.from_numpy(scan).to(dev)

            def c2_step():
                ctx.set_cloud_device(d_scan.data_ptr(), scan.shape[0], 3)
                ctx.step_normals_rsd(0.03, 0.03, ndiv=NDIV, plane_radius=PLANE_RADIUS)

            ms_c2 = timed_steps(c2_step, 10)
            pc2 = ctx.profile()
            k2 = pc2["neighbour_sum"] / scan.shape[0]
            b2 = scan.shape[0] * (48.0 * k2 + 192.0)
            extras["c2_scan_1M_r3cm"] = {"points_per_s": scan.shape[0] / (ms_c2 * 1e-3), "ms_per_step": ms_c2, "mean_neighbours": k2,
                                         "phases_ms": {"build": pc2["build_ms"], "normals": pc2["normals_ms"], "rsd": pc2["rsd_ms"]},
                                         "roofline_frac_whole_step": b2 / (ms_c2 * 1e-3) / 1e9 / peak}
            if not args.no_cpu_baseline:
                sub = slab_sample(scan, 100_000)
                rate2, dt2c, cores2 = cpu_points_per_s(sub, radius=0.03)
                pg = parity_gate(ctx, ctx_x if not args.exact else None, sub, radius=0.03, sens_points=sub.shape[0])
                extras["c2_scan_1M_r3cm"]["cpu_baseline"] = {"value": rate2, "unit": UNIT, "cores": cores2, "kind": "port",
                                                             "sample": f"x-slab of the scan, {sub.shape[0]} points"}
                extras["c2_scan_1M_r3cm"]["parity"] = pg
            del d_scan
        except Exception as e:
            extras["c2_scan_1M_r3cm"] = {"error": repr(e)}
        try:  # C3: GRSD clouds/s, the second half of BASELINE.json's metric
            g_ctx = ctx_x if not args.exact else ctx
            extras["c3_grsd_512"], _ = grsd_c3(g_ctx, args.clusters, 5, cpu_check=0 if args.no_cpu_baseline else 32)
        except Exception as e:
            extras["c3_grsd_512"] = {"error": repr(e)}

        # ---- CPU baseline (oracle port) + the parity gates of this line ---------------------------------------
        if not args.no_cpu_baseline:
            sample = slab_sample(pts, args.cpu_sample)
            rate, dt, cores = cpu_points_per_s(sample)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
                   "sample": f"x-slab of the room cloud, {sample.shape[0]} points (same density), oracle normals+RSD streaming mode"}
            try:
                parity = parity_gate(ctx if not args.exact else ctx_x, ctx_x if not args.exact else ctx, sample)
            except Exception as e:  # a failed gate must show up in the line, not kill the measurement
                parity = {"error": repr(e)}
            try:
                cpu["c1_reference_faithful"] = c1_reference_faithful(ctx)
            except Exception as e:
                cpu["c1_reference_faithful"] = {"error": repr(e)}
        ctx_x.close()

    if rank == 0:
        par = "single GPU"
        if world > 1:
            par = (f"query-shard x{world}: cloud replicated, every rank keys / sorts / tabulates only its slab of rows (+ 1 layer), the halo rows' "
                   "normals exchanged between neighbouring ranks over NVLink peer memory, results CONCATENATED inside the timed step: the RSD "
                   "kernel stores every point's normal and radii into the arrays of the rank that owns its input index (input order; rank g "
                   "holds input indices [n g / N, n (g + 1) / N)); cuts from a cost model + measured-time feedback; one C call per step per rank")
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if not args.exact else "f64", "data": "synthetic",
            "config": {"workload": "C4 20M-point synthetic room, normals+RSD r=2cm, query-sharded", "points": n,
                       "radius_m": RADIUS, "distance_div": NDIV, "plane_radius": PLANE_RADIUS, "max_nn": "unlimited",
                       "mean_neighbours": kbar, "candidates_tested_per_query": candidate_sum / max(1.0, n),
                       "l2": "inputs_larger_than_l2 (pos+normals 640 MB vs 126 MB L2)", "parallelism": par,
                       "mode": "exact-fp64" if args.exact else "fast-fp32", "kernels": KERNELS_VERSION,
                       "host_affinity": (f"rank 0 bound to the {numa_cores} cores of its GPU's NUMA node" if numa_cores else "not bound")},
            "phases_ms": {"build": build_ms, "normals": nrm_ms, "rsd": rsd_ms, "exchange_tail": my_phase["exchange_ms"],
                          "step_on_device": my_phase["step_ms"]},
            "per_rank_phase_ms": per_rank, "results_concatenated": concatenated,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        out.update(extras)
        if "other_mode" in extras and "ms_per_step" in extras["other_mode"]:
            out["exact_ms_per_step" if not args.exact else "fast_ms_per_step"] = extras["other_mode"]["ms_per_step"]
        if "c3_grsd_512" in extras and "clouds_per_s" in extras["c3_grsd_512"]:
            out["grsd_clouds_per_s"] = extras["c3_grsd_512"]["clouds_per_s"]
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
