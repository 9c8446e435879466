"""PCD I/O of the GRSD tools (second half) and the compute_grsd command line tool:
the reader on ASCII, binary v.7 and page-padded binary files, the writer's exact format, and -- on the GPU -- the
tool end to end against the oracle on the reference's shape_data clouds."""
import pathlib
import subprocess

import numpy as np
import pytest

from mapping_private_b200 import plugin, synth

ROOT = pathlib.Path(__file__).resolve().parent.parent
TOOL = ROOT / "mapping-private_b200" / "host" / "compute_grsd"


@pytest.fixture(scope="module")
def built():
    from mapping_private_b200 import cab

    cab.build()
    plugin.build()
    return plugin.LIB_PATH


def _write_pcd(path, xyz, normals=None, mode="binary", padded=False, extra_int=False):
    n = len(xyz)
    fields, size, typ, cols = ["x", "y", "z"], ["4"] * 3, ["F"] * 3, [xyz.astype("<f4")]
    if extra_int:  # an unrelated int32 column between the coordinates and the normals
        fields.append("idx"); size.append("4"); typ.append("I"); cols.append(np.arange(n, dtype="<i4")[:, None].view("<f4"))
    if normals is not None:
        fields += ["normal_x", "normal_y", "normal_z"]; size += ["4"] * 3; typ += ["F"] * 3; cols.append(normals.astype("<f4"))
    head = (f"# .PCD v.7 - Point Cloud Data file format\nVERSION .7\nFIELDS {' '.join(fields)}\nSIZE {' '.join(size)}\n"
            f"TYPE {' '.join(typ)}\nCOUNT {' '.join(['1'] * len(fields))}\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\n"
            f"POINTS {n}\nDATA {mode}\n")
    rec = np.concatenate(cols, axis=1)
    with open(path, "wb") as f:
        f.write(head.encode())
        if mode == "ascii":
            for i, row in enumerate(rec):
                vals = [repr(float(v)) for v in row]
                if extra_int:
                    vals[3] = str(i)
                f.write((" ".join(vals) + "\n").encode())
        else:
            if padded:
                f.write(b"\0" * (4096 - len(head)))
            f.write(rec.tobytes())


def test_pcd_reader_formats(built, tmp_path):
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(257, 3)).astype(np.float32)
    nrm = rng.normal(size=(257, 3)).astype(np.float32)
    xyz[5, 1] = np.nan
    for mode, padded, with_n, extra in [("ascii", False, True, True), ("binary", False, True, True), ("binary", True, False, False),
                                        ("binary", True, True, False)]:
        p = tmp_path / f"c_{mode}_{padded}_{with_n}.pcd"
        _write_pcd(p, xyz, nrm if with_n else None, mode, padded, extra)
        got, gn = plugin.pcd_read(p)
        assert np.array_equal(got.view(np.uint32), xyz.view(np.uint32)), (mode, padded)
        assert (gn is None) == (not with_n)
        if with_n:
            assert np.array_equal(gn, nrm)
    with pytest.raises(IOError):
        plugin.pcd_read(tmp_path / "missing.pcd")
    (tmp_path / "junk.pcd").write_text("hello\n")
    with pytest.raises(IOError):
        plugin.pcd_read(tmp_path / "junk.pcd")


def test_feature_writer_format(built, tmp_path):
    feat = np.array([[1.5, 0, 2], [0, 0, 0], [3, 4, 5.25]], np.float32)
    plugin.write_feature(tmp_path / "f.pcd", feat, remove_0=True)
    txt = (tmp_path / "f.pcd").read_text().splitlines()
    # grsd_colorCHLAC_tools.hpp:41-56
    assert txt[:9] == ["# .PCD v.7 - Point Cloud Data file format", "FIELDS vfh", "SIZE 4", "TYPE F", "COUNT 3", "WIDTH 2", "HEIGHT 1",
                       "POINTS 2", "DATA ascii"]
    assert txt[9] == "1.500000 0.000000 2.000000 " and txt[10] == "3.000000 4.000000 5.250000 " and len(txt) == 11
    plugin.write_feature(tmp_path / "g.pcd", feat, remove_0=False)
    assert (tmp_path / "g.pcd").read_text().splitlines()[5] == "WIDTH 3"


@pytest.mark.gpu
def test_compute_grsd_tool_against_oracle(built, oracle, kat, tmp_path):
    """The reference's noiseless sphere and cone clouds (page-padded binary PCD, as shipped) through the tool."""
    for shape, leaf in (("sphere", 0.01), ("cone", 0.01)):
        xyz = kat[f"{shape}_xyz"]
        src = tmp_path / f"{shape}.pcd"
        _write_pcd(src, xyz, None, "binary", padded=True)
        out = tmp_path / f"{shape}_grsd.pcd"
        r = subprocess.run([str(TOOL), str(src), str(leaf), str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        vals = np.loadtxt(out, skiprows=9).reshape(-1, 20)
        want = oracle.grsd_cluster(xyz, leaf, oracle.SIG_GRSD21)
        assert np.array_equal(vals[0].astype(np.int64), want["hist"][0][:20])
        # sliding boxes over every offset, PlusGRSD-110, all-zero histograms dropped by the writer
        r = subprocess.run([str(TOOL), str(src), str(leaf), str(out), "-subdiv", "4", "-offset", "2", "-kind", "110"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        vals = np.loadtxt(out, skiprows=9).reshape(-1, 110)
        rows = []
        for ox in (0, 2):
            for oy in (0, 2):
                for oz in (0, 2):
                    w = oracle.grsd_cluster(xyz, leaf, oracle.SIG_PLUSGRSD110, 4, (ox, oy, oz))
                    rows += [h for h in w["hist"] if h.any()]
        assert np.array_equal(vals.astype(np.int64), np.array(rows))
    r = subprocess.run([str(TOOL), str(tmp_path / "nope.pcd"), "0.01", str(tmp_path / "x.pcd")], capture_output=True, text=True)
    assert r.returncode != 0 and "Couldn't read file" in r.stderr
