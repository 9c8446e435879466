"""Small run of the step for compute-sanitizer (memcheck / racecheck)."""
import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pkgpath; pkgpath.load()
from mapping_private_b200 import cab, synth
pts = synth.tabletop(int(sys.argv[1]) if len(sys.argv) > 1 else 20000, noise_sigma=0.0003)
pts[11] = np.nan
c = cab.Context(0)
c.upload(pts)
res = []
for _ in range(3):
    c.step_normals_rsd(0.02, 0.02)
    res.append(c.download())
for r in res[1:]:
    print("same", all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(r, res[0])))
print(c.profile()["neighbour_sum"])
