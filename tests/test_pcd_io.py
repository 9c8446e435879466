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


def _write_pcd(path, xyz, normals=None, mode="binary", padded=False, extra_int=False, rgb=None):
    n = len(xyz)
    fields, size, typ, cols = ["x", "y", "z"], ["4"] * 3, ["F"] * 3, [xyz.astype("<f4")]
    if rgb is not None:  # PCL's packed colour: a float field whose bits are 0x00RRGGBB (the shipped shape_data layout: x y z rgb)
        fields.append("rgb"); size.append("4"); typ.append("F"); cols.append(np.asarray(rgb, "<u4")[:, None].view("<f4"))
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


def test_pcd_reader_rgb(built, tmp_path):
    rng = np.random.default_rng(1)
    xyz = rng.normal(size=(100, 3)).astype(np.float32)
    rgb = (rng.integers(1, 256, 100).astype(np.uint32) << 16) | (rng.integers(0, 256, 100).astype(np.uint32) << 8) | 7
    for mode, padded in (("binary", True), ("binary", False), ("ascii", False)):
        p = tmp_path / f"rgb_{mode}_{padded}.pcd"
        _write_pcd(p, xyz, None, mode, padded, rgb=rgb)
        assert np.array_equal(plugin.pcd_read(p)[0].view(np.uint32), xyz.view(np.uint32))
        assert np.array_equal(plugin.pcd_read_rgb(p), rgb), mode
    _write_pcd(tmp_path / "plain.pcd", xyz)
    assert plugin.pcd_read_rgb(tmp_path / "plain.pcd") is None


@pytest.mark.gpu
def test_compute_grsd_tool_colour_modes(built, oracle, tmp_path):
    """exampleVOSCH.cpp / example_GRSD_CCHLAC.cpp on the reference's shipped cube and sphere clouds (x y z rgb, page-padded
    binary): the colour part against the reference's own *_GRSD_CCHLAC.pcd values, the VOSCH vector against the oracle."""
    d = np.load(ROOT / "tests" / "golden" / "shape_data_vosch.npz")
    for shape, color in (("cube", "purple"), ("sphere", "orange")):
        xyz = d[f"noiseless_{shape}_xyz"]
        rgb = np.full(len(xyz), d[f"noiseless_{shape}_{color}_rgb"], np.uint32)
        src = tmp_path / f"{shape}_{color}.pcd"
        _write_pcd(src, xyz, None, "binary", padded=True, rgb=rgb)
        out = tmp_path / "out.pcd"
        r = subprocess.run([str(TOOL), str(src), "0.01", str(out), "-kind", "cchlac"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        vals = np.loadtxt(out, skiprows=9).reshape(-1, 117)[0]
        gold = d[f"noiseless_{shape}_{color}_vosch137"][20:].astype(np.float64)
        vals[:6] *= 0.5  # the shipped files' revision halved the 0th-order bins
        vals[63:69] *= 0.5
        assert np.all(np.abs(vals - gold) <= 1.1e-6 * np.maximum(1.0, np.abs(gold)))  # both sides printed with %f
        r = subprocess.run([str(TOOL), str(src), "0.01", str(out), "-kind", "vosch"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        vosch = np.loadtxt(out, skiprows=9).reshape(-1, 137)[0]
        want = oracle.grsd_cluster(xyz, 0.01, oracle.SIG_GRSD21)
        assert np.array_equal(vosch[:20].astype(np.int64), want["hist"][0][:20])
        grid = oracle.voxel_grid(xyz, 0.01)
        grid["leaf"] = 0.01
        _, _, h = oracle.color_chlac117(grid, oracle.voxel_colors(xyz, rgb, 0.01), c3=True)
        assert np.all(np.abs(vosch[20:] - h[0].astype(np.float64)) <= 5.1e-7 * np.maximum(1.0, np.abs(h[0])))
    _write_pcd(tmp_path / "nocolor.pcd", d["noiseless_sphere_xyz"])
    r = subprocess.run([str(TOOL), str(tmp_path / "nocolor.pcd"), "0.01", str(tmp_path / "x.pcd"), "-kind", "vosch"], capture_output=True, text=True)
    assert r.returncode != 0 and "no rgb field" in r.stderr
