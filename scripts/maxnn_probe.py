"""Cost of the max_nn truncation path (the plugin defaults: max_nn 150, 75 in launch/pipeline_tmp.launch:20) on the
room cloud: threshold selection + thresholded RSD against the unlimited pass."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pkgpath; pkgpath.load()
from mapping_private_b200 import cab, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
pts = synth.room(n)
ctx = cab.Context(0)
ctx.upload(pts)
ctx.build_grid(0.02)
ctx.normals(0.02, download=False)
for max_nn in (0, 150, 75):
    for _ in range(2):
        t0 = time.perf_counter()
        ctx.rsd(0.02, max_nn=max_nn, download=False)
        dt = time.perf_counter() - t0
    p = ctx.profile()
    print(f"n {n} max_nn {max_nn:4d}: cab_rsd wall {dt * 1e3:8.2f} ms (rsd kernel {p['rsd_ms']:.2f} ms), k/q {p['neighbour_sum'] / n:.1f}")
