"""Extracts the known-answer fixtures the reference ships for the GRSD path.

Run in the build container (needs /root/reference); the outputs are committed so
that tests never read /root/reference at run time.

Inputs : color_chlac/demos/shape_data/noiseless_{shape}_blue.pcd  (PCD v.7, FIELDS x y z rgb,
         DATA binary, payload at byte offset 4096, N x 16 B little-endian float32)
Goldens: the matching *_GRSD_CCHLAC.pcd (ASCII, COUNT 137): the first 20 values are
         integer counts x 5/104 written by an OLDER revision of the code:
         slots (i, j), i = 1..5, j = 0..i, 13 half-stencil offsets, leaf 0.01.  For
         single-class shapes slot (c,0) = sum over voxels of empty half-neighbours and
         slot (c,c) = occupied half-neighbour pairs, total 13 * V.
"""
import pathlib
import re

import numpy as np

REF = pathlib.Path("/root/reference/color_chlac/demos/shape_data")
OUT = pathlib.Path(__file__).resolve().parent
SHAPES = ["plane", "sphere", "cylinder", "torus", "cone"]


def read_pcd_xyz(path):
    raw = path.read_bytes()
    head = raw[:4096].decode("ascii", errors="ignore")
    n = int(re.search(r"POINTS (\d+)", head).group(1))
    assert "DATA binary" in head
    start = head.index("DATA binary") + len("DATA binary\n")
    # header is page-padded: payload starts at 4096 when the header is shorter
    if len(raw) - 4096 == n * 16:
        start = 4096
    a = np.frombuffer(raw, dtype="<f4", count=n * 4, offset=start).reshape(n, 4)
    return np.ascontiguousarray(a[:, :3])


def svm_fixture():
    """The SVM model and scale ranges the reference pipeline loads for GRSD
    (dyn_obj_store/table_pipeline_grsd.launch: svm/grsd_ijrr.model + .scp), parsed with the repo's own
    parser and stored as arrays (libsvm model files are data, not source)."""
    import sys

    sys.path.insert(0, str(OUT.parent.parent))
    import pkgpath

    pkgpath.load()
    from mapping_private_b200 import svm_model

    svm = pathlib.Path("/root/reference/cloud_algos/svm")
    m = svm_model.parse_model((svm / "grsd_ijrr.model").read_text(), 21)
    lower, upper, fmin, fmax = svm_model.parse_scale((svm / "grsd_ijrr.scp").read_text(), 21)
    np.savez_compressed(OUT / "svm_grsd_ijrr.npz", gamma=m.gamma, labels=m.labels, nr_sv=m.nr_sv, rho=m.rho,
                        sv_coef=m.sv_coef, sv=m.sv, lower=lower, upper=upper, fmin=fmin, fmax=fmax)
    print("svm_grsd_ijrr:", m.nr_class, "classes,", m.total_sv, "SVs, dim", m.dim)


def vosch_fixture():
    """The reference's own VOSCH outputs (20 GRSD + 117 C3-HLAC values per file, exampleVOSCH.cpp: leaf 0.01, colour
    thresholds 127) for ten of its fourteen shape geometries in all seven colours.  Every file is uniformly coloured and
    the seven colours of a shape share one geometry, so the fixture holds each geometry once."""
    geoms = ["noiseless_" + s for s in ("plane", "sphere", "cylinder", "torus", "cone", "cube", "dice")] + \
            ["noisy_" + s for s in ("torus", "cone", "sphere")]
    colors = ["black", "blue", "green", "orange", "purple", "red", "yellow"]
    data = {"geoms": np.array(geoms), "colors": np.array(colors)}
    for g in geoms:
        xyz0 = None
        for c in colors:
            raw = (REF / f"{g}_{c}.pcd").read_bytes()
            head = raw[:4096].decode("ascii", errors="ignore")
            n = int(re.search(r"POINTS (\d+)", head).group(1))
            a = np.frombuffer(raw, dtype="<f4", count=n * 4, offset=4096).reshape(n, 4)
            xyz, rgb = np.ascontiguousarray(a[:, :3]), np.ascontiguousarray(a[:, 3]).view(np.uint32)
            assert len(np.unique(rgb)) == 1
            if xyz0 is None:
                xyz0 = xyz
                data[f"{g}_xyz"] = xyz.astype(np.float32)
            assert np.array_equal(xyz0, xyz)
            data[f"{g}_{c}_rgb"] = np.uint32(rgb[0])
            data[f"{g}_{c}_vosch137"] = np.loadtxt(REF / f"{g}_{c}_GRSD_CCHLAC.pcd", skiprows=9).astype(np.float32)
    np.savez_compressed(OUT / "shape_data_vosch.npz", **data)
    print("shape_data_vosch:", len(geoms), "geometries x", len(colors), "colours")


def normals_fixture():
    """color_chlac/demos/data/tmp_normal.pcd: 4712 points with the normal_x/y/z and curvature fields the reference's
    computeNormal (pcl::NormalEstimation, setRadiusSearch(0.02), grsd_colorCHLAC_tools.hpp:67-90) wrote for exactly these
    points -- recognised by reproducing them: radius 0.02 on the file's own points gives the stored normals to 2e-5 rad
    and the stored curvatures to 4e-7 (no other radius or k comes close)."""
    raw = (REF.parent / "data" / "tmp_normal.pcd").read_bytes()
    head = raw[:4096].decode("ascii", errors="ignore")
    fields = re.search(r"FIELDS (.*)", head).group(1).split()
    n = int(re.search(r"POINTS (\d+)", head).group(1))
    assert len(raw) - 4096 == n * 4 * len(fields)
    a = np.frombuffer(raw, dtype="<f4", count=n * len(fields), offset=4096).reshape(n, len(fields))
    col = {f: i for i, f in enumerate(fields)}
    np.savez_compressed(OUT / "tmp_normal.npz", xyz=np.ascontiguousarray(a[:, [col["x"], col["y"], col["z"]]]),
                        normal=np.ascontiguousarray(a[:, [col["normal_x"], col["normal_y"], col["normal_z"]]]),
                        curvature=np.ascontiguousarray(a[:, col["curvature"]]))
    print("tmp_normal:", n, "points")


def main():
    svm_fixture()
    normals_fixture()
    vosch_fixture()
    data = {}
    for s in SHAPES:
        xyz = read_pcd_xyz(REF / f"noiseless_{s}_blue.pcd")
        gold = np.loadtxt(REF / f"noiseless_{s}_blue_GRSD_CCHLAC.pcd", skiprows=9)[:20]
        counts = gold / (5.0 / 104.0)
        assert np.allclose(counts, np.rint(counts), atol=1e-2), s
        data[f"{s}_xyz"] = xyz.astype(np.float32)
        data[f"{s}_counts20"] = np.rint(counts).astype(np.int64)
        print(s, xyz.shape, np.rint(counts).astype(int)[np.rint(counts) != 0])
    np.savez_compressed(OUT / "shape_data_kat.npz", **data)


if __name__ == "__main__":
    main()
