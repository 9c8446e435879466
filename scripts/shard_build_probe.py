"""One shard's step under ncu: which kernels make up a sharded cab_step_normals_rsd (slab build + normals + RSD).
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/shard_build_probe.py [rank] [world]"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pkgpath; pkgpath.load()
from mapping_private_b200 import cab, synth

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 3
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pts = synth.room(20_000_000)
ctx = cab.Context(0)
ctx.upload(pts)
ctx.set_shard(rank, world)
for _ in range(3):
    ctx.step_normals_rsd(0.02, 0.02)
print(ctx.profile())
