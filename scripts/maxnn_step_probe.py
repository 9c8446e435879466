"""One step with the plugin default max_nn = 150 on the RSD pass (normals unlimited) under ncu: which kernels make it up.
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/maxnn_step_probe.py [points] [max_nn]"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pkgpath; pkgpath.load()
from mapping_private_b200 import cab, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
max_nn = int(sys.argv[2]) if len(sys.argv) > 2 else 150
pts = synth.room(n)
ctx = cab.Context(0)
ctx.upload(pts)
for _ in range(2):
    ctx.step_normals_rsd(0.02, 0.02, max_nn_rsd=max_nn)
print(ctx.profile())
