"""One density of the C5 sweep for an ncu pass (DRAM bytes and L2 hit rate of the two pass kernels):
  ncu --kernel-name regex:'normals_kernel|rsd_fast_kernel' --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
      --clock-control none --csv --log-file out.csv python scripts/density_ncu.py K [points]"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pkgpath; pkgpath.load()
from mapping_private_b200 import cab, synth

k = float(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
pts = synth.room(n) if k == 0 else synth.density_patches(n, k, 0.02)  # K = 0: the C4 room cloud
ctx = cab.Context(0)
ctx.upload(pts)
for _ in range(2):
    ctx.build_grid(0.02)
    ctx.normals(0.02, download=False)
    ctx.rsd(0.02, download=False)
p = ctx.profile()
print({"k_target": k, "points": pts.shape[0], "mean_neighbours": p["neighbour_sum"] / pts.shape[0], "candidates_per_query": p["candidate_sum"] / pts.shape[0]})
