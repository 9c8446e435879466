"""BASELINE config C5: density sweep on a 5 M-point cloud (mean 16-512 neighbours per query at
r = 2 cm), normals + RSD, kernel times and the achieved algorithmic GB/s against the measured HBM
peak (SURVEY section 8(d) accounting: normals 16 k + 32, RSD 32 k + 40, build 120 bytes per point).

  python scripts/density_sweep.py [--points N] [--out profiles/rNN_density_sweep.json]

Prints one JSON document.  Parity of the neighbour sets on these clouds is covered by
tests/test_gpu_parity.py::test_density_patches_neighbour_sets (the oracle is test infrastructure only)."""
import argparse
import json
import pathlib
import statistics
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pkgpath  # noqa: E402

pkgpath.load()
from mapping_private_b200 import cab, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=5_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    peak = 6650.0
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    r = 0.02
    ctx = cab.Context(0)
    rows = []
    for k_mean in (16, 32, 64, 128, 256, 512):
        pts = synth.density_patches(args.points, float(k_mean), r)
        n = pts.shape[0]
        ctx.upload(pts)
        t = {"build_ms": [], "normals_ms": [], "rsd_ms": []}
        for rep in range(args.reps + 2):
            ctx.build_grid(r)
            ctx.normals(r, download=False)
            ctx.rsd(r, download=False)
            if rep >= 2:
                pr = ctx.profile()
                for key in t:
                    t[key].append(pr[key])
        pr = ctx.profile()
        k = pr["neighbour_sum"] / n
        ms = {key: statistics.mean(v) for key, v in t.items()}
        nb, rb, bb = n * (16.0 * k + 32), n * (32.0 * k + 40), n * 120.0
        row = {"k_target": k_mean, "points": n, "mean_neighbours": k, "candidates_per_query": pr["candidate_sum"] / n,
               **ms, "step_ms": sum(ms.values()), "points_per_s": n / (sum(ms.values()) * 1e-3),
               "normals_GBs": nb / ms["normals_ms"] / 1e6, "rsd_GBs": rb / ms["rsd_ms"] / 1e6, "build_GBs": bb / ms["build_ms"] / 1e6,
               "step_GBs": (nb + rb + bb) / sum(ms.values()) / 1e6}
        row["step_frac_of_hbm_peak"] = row["step_GBs"] / peak
        row["rsd_frac_of_hbm_peak"] = row["rsd_GBs"] / peak
        rows.append(row)
        print(json.dumps(row), file=sys.stderr)
    doc = {"config": "C5 density sweep, normals + RSD r = 2 cm, 1 x B200, fast-fp32 mode", "hbm_peak_GBs": peak, "rows": rows}
    text = json.dumps(doc, indent=1)
    if args.out:
        pathlib.Path(args.out).write_text(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
