"""Euclidean clustering timings (cab_euclidean_clusters): the objects of the tabletop scene, the 1 M-point scan
and the 20 M-point room (one giant component: the worst case for the size / minimum-index atomics)."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pkgpath; pkgpath.load()
import numpy as np
from mapping_private_b200 import cab, synth

ctx = cab.Context(0)
cases = [("tabletop objects 400k, tol 2 cm", lambda: (lambda p: p[p[:, 2] > 0.754])(synth.tabletop(1_000_000)), 0.02, 30),
         ("scan 1M, tol 3 cm", lambda: synth.scan(1_000_000), 0.03, 30),
         ("scan 1M, tol 5 cm", lambda: synth.scan(1_000_000), 0.05, 30),
         ("room 20M, tol 1 cm", lambda: synth.room(20_000_000), 0.01, 30),
         ("room 20M, tol 2 cm", lambda: synth.room(20_000_000), 0.02, 30)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for name, gen, tol, min_pts in cases:
    pts = np.ascontiguousarray(gen(), np.float32)
    ctx.upload(pts)
    for _ in range(2):
        t0 = time.perf_counter()
        lab, nc = ctx.euclidean_clusters(tol, min_pts)
        wall = time.perf_counter() - t0
    p = ctx.profile()
    sizes = np.bincount(lab[lab >= 0]) if nc else np.zeros(1, int)
    print(f"{name}: n {len(pts)} clusters {nc} largest {sizes.max()} unlabelled {(lab < 0).sum()} build {p['build_ms']:.2f} ms cluster {p['cluster_ms']:.2f} ms wall {1e3 * wall:.1f} ms")
