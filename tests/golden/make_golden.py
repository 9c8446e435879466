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


def main():
    svm_fixture()
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
