"""Runs the W shards of the 20M-point room one after the other on ONE GPU and prints per-shard
kernel times and work statistics (used to tune the multi-GPU split without an 8-GPU box): each shard's
cab_step_normals_rsd (slab build + normals of own and halo rows + RSD of own rows), without the result exchange.
usage: python scripts/shard_probe.py [W] [points] [halo_permille]"""
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
    best = None
    for _ in range(4):
        ctx.step_normals_rsd(0.02, 0.02)
        p = ctx.profile()
        if best is None or p['step_ms'] < best['step_ms']:
            best = p
    b, e = ctx.shard_range()
    q = max(e - b, 1)
    rows.append((g, b, e, best))
tot = 0.0
for g, b, e, p in rows:
    q = max(e - b, 1)
    print(f"shard {g:3d}: q {q:8d} sorted {p['n_sorted']:9d} packets {p['n_packets']:7d} build {p['build_ms']:6.3f} normals {p['normals_ms']:6.3f} "
          f"rsd {p['rsd_ms']:6.3f} step {p['step_ms']:6.3f} ms  k/q {p['neighbour_sum'] / q:6.1f} cand/q {p['candidate_sum'] / q:7.1f}")
    tot = max(tot, p['step_ms'])
print(f"W {W}: slowest shard step {tot:.3f} ms")
