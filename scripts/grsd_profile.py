import sys, time, pathlib
sys.path.insert(0, "/root/repo")
import pkgpath; pkgpath.load()
import numpy as np
from mapping_private_b200 import cab, synth
xyz, off = synth.clusters(512)
for exact in (True, False):
    ctx = cab.Context(0, exact=exact)
    for i in range(4):
        t0 = time.perf_counter(); h = ctx.grsd_batch(xyz, off, 0.025, r_normals=0.02); dt = time.perf_counter() - t0
    p = ctx.profile()
    print("exact", exact, "wall ms", round(dt * 1e3, 2), {k: round(p[k], 3) for k in ("h2d_ms", "build_ms", "normals_ms", "grsd_ms")}, "voxels", len(ctx.grsd_voxels(512)["labels"]))
