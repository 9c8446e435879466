"""Debug: local group on one GPU, results after every step against the single-GPU results."""
import os, sys, threading
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import pkgpath
pkgpath.load()
from mapping_private_b200 import cab, synth

R = 0.02
world = int(sys.argv[1]) if len(sys.argv) > 1 else 3
max_nn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
layout = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pts = synth.tabletop(90_000, noise_sigma=0.0002)
pts[5] = np.nan
n = pts.shape[0]
c = cab.Context(0)
c.upload(pts); c.build_grid(R)
n4 = c.normals(R, max_nn=max_nn); rmin, rmax = c.rsd(R, max_nn=max_nn)
c.close()
ctxs = [cab.Context(0) for _ in range(world)]
cab.comm_init_local(ctxs)
bar = threading.Barrier(world)
lock = threading.Lock()
def same(a, b):
    return np.ascontiguousarray(a).view(np.uint32) == np.ascontiguousarray(b).view(np.uint32)
def work(r):
    c = ctxs[r]
    if hasattr(c, "comm_set_layout"): c.comm_set_layout(layout)
    for it in range(3):
        c.comm_upload_cloud(pts)
        c.step_normals_rsd(R, R, max_nn_normals=max_nn, max_nn_rsd=max_nn)
        bar.wait()
        lo, hi = (0, n) if layout == 0 else (n * r // world, n * (r + 1) // world)
        f4, fmin, fmax = c.comm_download_range(lo, hi)
        bad4 = ~same(f4, n4[lo:hi]).all(axis=1); badr = ~(same(fmin, rmin[lo:hi]) & same(fmax, rmax[lo:hi]))
        p = c.profile()
        with lock:
            print(f"it {it} rank {r} range {c.shard_range()} mode {p['shard_mode']} n_sorted {p['n_sorted']} bad normals {bad4.sum()} bad radii {badr.sum()}"
                  f" nan-in-result {np.isnan(f4[:,0]).sum()} first bad {np.flatnonzero(bad4)[:5]} ksum {p['neighbour_sum']}", flush=True)
        bar.wait()
ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
[t.start() for t in ts]; [t.join() for t in ts]
