"""Runs the W shards of the 20M-point room one after the other on ONE GPU and prints per-shard
kernel times and work statistics (used to tune the multi-GPU split without an 8-GPU box)."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pkgpath; pkgpath.load()
from mapping_private_b200 import cab, synth

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
pts = synth.room(n)
ctx = cab.Context(0)
ctx.upload(pts)
rows = []
for g in range(W):
    ctx.set_shard(g, W)
    ctx.build_grid(0.02)
    ctx.build_grid(0.02)
    pb = ctx.profile()
    b, e = ctx.shard_range()
    for _ in range(2):
        ctx.normals(0.02, download=False)
        p1 = ctx.profile()
        ctx.rsd(0.02, download=False)
        p2 = ctx.profile()
    q = max(e - b, 1)
    rows.append((g, b, e, p1['normals_ms'], p2['rsd_ms'], p2['neighbour_sum'] / q, p2['candidate_sum'] / q, pb['build_ms'], pb['n_sorted']))
for g, b, e, tn, tr, k, c, tb, ns in rows:
    q = max(e - b, 1)
    print(f"shard {g:3d}: [{b:9d},{e:9d}) q {q:8d} normals {tn:6.3f} ms rsd {tr:6.3f} ms k/q {k:6.1f} cand/q {c:7.1f} ns/q n {1e6*tn/q:6.2f} r {1e6*tr/q:6.2f} build {tb:6.3f} ms sorted {ns:9d} total {tb+tn+tr:6.3f} ms")
