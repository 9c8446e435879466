"""Pins the CPU oracle: brute-force / scipy cross-checks of the neighbour sets, analytic
known answers for normals and RSD radii, the get_type truth table, hand-built voxel blocks,
subdivision cases and the known answers decoded from the reference's shape_data goldens
(tests/golden/make_golden.py)."""
import pathlib

import numpy as np
import pytest
from scipy.spatial import cKDTree

from mapping_private_b200 import synth

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


def _sets(off, idx):
    return [idx[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def test_d2_rule(oracle):
    L = oracle.lib()
    import ctypes as C
    a = np.array([0.1, 0.2, 0.3], np.float32)
    b = np.array([0.11, 0.18, 0.33], np.float32)
    d = a - b
    want = np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2]))
    got = L.orc_d2(a.ctypes.data_as(C.POINTER(C.c_float)), b.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.float32(got) == want


@pytest.mark.parametrize("max_nn", [0, 20])
def test_radius_search_vs_brute(oracle, max_nn):
    p = synth.tabletop(6000)
    q = p[::7]
    o1, i1, d1 = oracle.radius_search(p, q, 0.03, max_nn=max_nn)
    o2, i2, d2 = oracle.radius_search(p, q, 0.03, max_nn=max_nn, brute=True)
    assert np.array_equal(o1, o2) and np.array_equal(i1, i2) and np.array_equal(d1, d2)
    # sorted by (d2, idx), self first
    for s, (a, b) in enumerate(zip(o1[:-1], o1[1:])):
        dd, ii = d1[a:b], i1[a:b]
        assert np.all(np.diff(dd) >= 0)
        assert dd[0] == 0.0


def test_radius_search_vs_scipy(oracle):
    # lattice coordinates make d2 exact, so a float64 k-d tree must agree bit-for-bit
    p = synth.tabletop(20000)
    r = 0.02
    r2 = float(np.float32(r) * np.float32(r))
    off, idx, d2 = oracle.radius_search(p, p, r)
    tree = cKDTree(p.astype(np.float64))
    lists = tree.query_ball_point(p.astype(np.float64), np.sqrt(r2) * (1 + 1e-9))
    for i in range(0, p.shape[0], 13):
        cand = np.array(lists[i])
        dd = np.sum((p[cand].astype(np.float64) - p[i].astype(np.float64)) ** 2, axis=1)
        want = np.sort(cand[dd <= r2])
        got = np.sort(idx[off[i]:off[i + 1]])
        assert np.array_equal(want, got)


def test_radius_search_edge_cases(oracle):
    empty = np.zeros((0, 3), np.float32)
    off, idx, _ = oracle.radius_search(empty, empty, 0.02)
    assert off.tolist() == [0] and idx.size == 0
    one = np.array([[1, 2, 3]], np.float32)
    off, idx, d2 = oracle.radius_search(one, one, 0.02)
    assert off.tolist() == [0, 1] and idx.tolist() == [0] and d2.tolist() == [0.0]
    # duplicates and a point exactly at the radius (inclusive rule d2 <= r2)
    r = np.float32(0.0625)
    p = np.array([[0, 0, 0], [0, 0, 0], [r, 0, 0], [np.nextafter(r, np.float32(1)), 0, 0], [np.nan, 0, 0]], np.float32)
    off, idx, d2 = oracle.radius_search(p, p[:1], float(r))
    assert idx.tolist() == [0, 1, 2]


def test_normals_analytic(oracle):
    plane = synth.analytic_shape("plane", 8000)
    n4, k = oracle.normals(plane, 0.02)
    assert np.all(np.abs(np.abs(n4[:, 2]) - 1) < 1e-6) and np.all(n4[:, 3] < 1e-6)
    # viewpoint (0,0,0) is below the plane z=1 -> normals point down
    assert np.all(n4[:, 2] < 0)
    sph = synth.analytic_shape("sphere", 20000, R=0.05)
    n4, k = oracle.normals(sph, 0.01)
    radial = (sph - np.array([0.5, 0.5, 1.0], np.float32)) / 0.05
    cosang = np.abs(np.sum(radial * n4[:, :3], axis=1))
    assert np.percentile(np.degrees(np.arccos(np.clip(cosang, 0, 1))), 99) < 3.0
    assert np.allclose(np.linalg.norm(n4[:, :3], axis=1), 1, atol=1e-6)
    # fewer than 3 neighbours -> NaN
    lone = np.array([[0, 0, 0], [1, 0, 0], [1, 0.001, 0]], np.float32)
    n4, k = oracle.normals(lone, 0.02)
    assert k.tolist() == [1, 2, 2] and np.all(np.isnan(n4))


def test_normals_vs_numpy_pca(oracle):
    p = synth.tabletop(5000, noise_sigma=0.0005)
    r = 0.03
    n4, k = oracle.normals(p, r)
    off, idx, _ = oracle.radius_search(p, p, r)
    for i in range(0, 5000, 97):
        nb = p[idx[off[i]:off[i + 1]]].astype(np.float64)
        assert nb.shape[0] == k[i]
        w, v = np.linalg.eigh(np.cov(nb.T, bias=True))
        n = v[:, 0]
        if (w[1] - w[0]) / w.sum() < 1e-6:
            continue
        ang = np.linalg.norm(np.cross(n, n4[i, :3].astype(np.float64)))  # sin(angle), sign-insensitive
        assert ang < 2e-6
        assert abs(w[0] / w.sum() - n4[i, 3]) < 1e-6


def test_rsd_analytic(oracle):
    plane = synth.analytic_shape("plane", 8000)
    n4, _ = oracle.normals(plane, 0.02)
    rmin, rmax, rdif = oracle.rsd(plane, n4, 0.02)
    assert np.all(rmin == np.float32(0.1)) and np.all(rmax == np.float32(0.1)) and np.all(rdif == 0)
    sph = synth.analytic_shape("sphere", 20000, R=0.05)
    n4, _ = oracle.normals(sph, 0.02)
    rmin, rmax, _ = oracle.rsd(sph, n4, 0.02)
    assert abs(np.median(rmin) - 0.05) < 0.0075 and abs(np.median(rmax) - 0.05) < 0.0075
    cyl = synth.analytic_shape("cylinder", 20000, R=0.04)
    n4, _ = oracle.normals(cyl, 0.02)
    rmin, rmax, _ = oracle.rsd(cyl, n4, 0.02)
    inner = np.abs(cyl[:, 2] - 1.0) < 0.12
    assert abs(np.median(rmin[inner]) - 0.04) < 0.006
    assert np.median(rmax[inner]) == np.float32(0.1)


def test_rsd_matches_python_restatement(oracle):
    """Independent pure-Python restatement of radius_estimation.cpp:140-202 on a small cloud."""
    p = synth.tabletop(3000, noise_sigma=0.0003)
    r, ndiv, plane_r = 0.03, 10, 0.1
    n4, _ = oracle.normals(p, r)
    for max_nn in (0, 25):
        rmin, rmax, rdif = oracle.rsd(p, n4, r, max_nn=max_nn, ndiv=ndiv, plane_radius=plane_r)
        off, idx, d2 = oracle.radius_search(p, p, r, max_nn=max_nn)
        nf = n4[:, :3]
        for cp in range(0, 3000, 41):
            mn = [np.inf] * ndiv
            mx = [-np.inf] * ndiv
            for s in range(off[cp], off[cp + 1]):
                j = idx[s]
                if j == cp:
                    continue
                c = np.float32(np.float32(np.float32(nf[cp, 0] * nf[j, 0]) + np.float32(nf[cp, 1] * nf[j, 1])) + np.float32(nf[cp, 2] * nf[j, 2]))
                c = min(1.0, max(-1.0, float(c)))
                ang = np.arccos(c)
                if ang > np.pi / 2:
                    ang = np.pi - ang
                b = min(ndiv - 1, int(np.floor(ndiv * float(np.sqrt(np.float32(d2[s]))) / r)))  # std::sqrt(float), :165
                mn[b] = min(mn[b], ang)
                mx[b] = max(mx[b], ang)
            a_nn = a_nd = a_xx = a_xd = 0.0
            for di in range(ndiv):
                if mx[di] >= 0:
                    f = (di + 0.5) * r / ndiv
                    a_nn += mn[di] * mn[di]
                    a_nd += mn[di] * f
                    a_xx += mx[di] * mx[di]
                    a_xd += mx[di] * f
            want_max = plane_r if a_nn == 0 else min(a_nd / a_nn, plane_r)
            want_min = plane_r if a_xx == 0 else min(a_xd / a_xx, plane_r)
            assert abs(rmin[cp] - want_min) <= 1e-6 * want_min
            assert abs(rmax[cp] - want_max) <= 1e-6 * want_max
            assert abs(rdif[cp] - (want_max - want_min)) <= 1e-6


def test_rsd_degenerate_inputs(oracle):
    # isolated point and NaN normals -> plane_radius
    p = np.array([[0, 0, 0], [1, 0, 0], [1.001, 0, 0]], np.float32)
    n = np.array([[0, 0, 1], [np.nan, 0, 0], [0, 0, 1]], np.float32)
    rmin, rmax, rdif = oracle.rsd(p, n, 0.02, plane_radius=0.1)
    assert np.all(rmin == np.float32(0.1)) and np.all(rmax == np.float32(0.1))
    # identical normals -> all angles 0 -> zero denominators -> plane_radius
    q = synth.analytic_shape("plane", 500)
    nn = np.tile(np.array([[0, 0, 1]], np.float32), (500, 1))
    rmin, rmax, _ = oracle.rsd(q, nn, 0.02)
    assert np.all(rmin == np.float32(0.1))


def test_ref_faithful_equals_streaming(oracle):
    p = synth.tabletop(4000)
    n4, _ = oracle.normals(p, 0.02)
    for max_nn in (0, 30):
        a, b, _ = oracle.rsd(p, n4, 0.02, max_nn=max_nn)
        c, d, ph = oracle.rsd_ref_faithful(p, n4, 0.02, max_nn=max_nn)
        assert np.array_equal(a, c) and np.array_equal(b, d) and np.all(ph >= 0)


def test_get_type_truth_table(oracle):
    f = np.float32
    up = lambda x: float(np.nextafter(f(x), f(10)))
    dn = lambda x: float(np.nextafter(f(x), f(-10)))
    # grsd_colorCHLAC_tools.hpp:104-116 (fp32 arguments against double literals)
    assert oracle.get_type(up(0.100), 0.2) == 1  # PLANE
    assert oracle.get_type(0.2, 0.2) == 1
    assert oracle.get_type(dn(0.100), up(0.175)) == 2  # CYLINDER
    assert oracle.get_type(0.05, 0.2) == 2
    assert oracle.get_type(dn(0.015), 0.1) == 0  # NOISE
    assert oracle.get_type(0.0, 0.0) == 0
    assert oracle.get_type(0.05, 0.055) == 3  # SPHERE
    assert oracle.get_type(0.02, 0.0699) == 3
    assert oracle.get_type(0.02, 0.0701) == 4  # EDGE
    assert oracle.get_type(0.05, 0.15) == 4
    # float(0.1) > 0.1 (double) because float(0.1) = 0.100000001490116
    assert oracle.get_type(float(f(0.1)), 0.2) == 1
    assert oracle.get_type(float(f(0.015)), float(f(0.015))) == 0  # float(0.015) < 0.015


def test_offsets26(oracle):
    o = oracle.offsets26()
    assert o[0].tolist() == [-1, -1, -1] and o[8].tolist() == [1, 1, -1]
    assert o[9].tolist() == [-1, -1, 0] and o[11].tolist() == [1, -1, 0] and o[12].tolist() == [-1, 0, 0]
    assert np.array_equal(o[13:], -o[:13])
    assert len({tuple(r) for r in o.tolist()}) == 26 and (0, 0, 0) not in {tuple(r) for r in o.tolist()}


def test_voxel_grid_semantics(oracle):
    leaf = 0.01
    p = np.array([[0.001, 0.001, 0.001], [0.009, 0.003, 0.005], [0.011, 0.001, 0.001], [-0.001, 0.0, 0.0]], np.float32)
    g = oracle.voxel_grid(p, leaf)
    assert g["min_b"].tolist() == [-1, 0, 0] and g["div_b"].tolist() == [3, 1, 1] and g["nvox"] == 3
    assert g["counts"].tolist() == [1, 2, 1]
    assert np.allclose(g["centroids"][1], [0.005, 0.002, 0.003], atol=1e-7)
    assert g["layout"].tolist() == [0, 1, 2]


def test_transitions_hand_block(oracle):
    """3x3x3 fully occupied block, centre voxel labelled SPHERE(3), rest PLANE(1)."""
    leaf = np.float32(0.01)
    ii, jj, kk = np.meshgrid(range(3), range(3), range(3), indexing="ij")
    pts = (np.stack([ii, jj, kk], -1).reshape(-1, 3) + 0.5) * 0.01
    pts = pts.astype(np.float32)
    g = oracle.voxel_grid(pts, float(leaf))
    g["leaf"] = float(leaf)
    assert g["nvox"] == 27
    types = np.ones(27, np.int32)
    centre = g["layout"][1 + 1 * 3 + 1 * 9]
    types[centre] = 3
    M, h = oracle.grsd_transitions(g, types)
    # per-voxel in-grid neighbours: 8 corners x7, 12 edges x11, 6 faces x17, centre x26
    occupied_pairs = 8 * 7 + 12 * 11 + 6 * 17 + 26
    assert M.sum() == 27 * 26
    assert M[3, 1] == 26 and M[1, 3] == 26 and M[3, 3] == 0 and M[3, 5] == 0
    assert M[1, 1] == occupied_pairs - 52
    assert M[1, 5] == 27 * 26 - occupied_pairs
    # packing: upper triangle, row-major (grsd_colorCHLAC_tools.hpp:271-275)
    want = [M[i, j] for i in range(6) for j in range(i, 6)]
    assert h.tolist() == want and h[20] == 0


def test_subdivision_cases(oracle):
    leaf = 0.01
    ii, jj, kk = np.meshgrid(range(4), range(4), range(2), indexing="ij")
    pts = ((np.stack([ii, jj, kk], -1).reshape(-1, 3) + 0.5) * 0.01).astype(np.float32)
    g = oracle.voxel_grid(pts, leaf)
    g["leaf"] = leaf
    types = np.ones(g["nvox"], np.int32)
    hn, sb, h = oracle.grsd21_subdiv(g, types, 2)
    assert sb.tolist() == [2, 2, 1] and hn == 4 and h.shape == (4, 21)
    _, h1 = oracle.grsd_transitions(g, types)
    assert np.array_equal(h.sum(0), h1)  # subdivisions partition the voxels
    assert np.all(h.sum(1) == 8 * 26)
    # offsets >= grid size -> zero vector (grsd_colorCHLAC_tools.hpp:150-153)
    hn, sb, h = oracle.grsd21_subdiv(g, types, 2, off=(4, 0, 0))
    assert hn == 0 and sb.tolist() == [0, 0, 0]
    # negative subdivision size is rejected (:158-161)
    hn, _, _ = oracle.grsd21_subdiv(g, types, -1)
    assert hn == -1
    # offset skips the voxels below it (:243)
    hn, sb, h = oracle.grsd21_subdiv(g, types, 2, off=(1, 0, 0))
    assert sb.tolist() == [2, 2, 1] and h.sum() == (3 * 4 * 2) * 26
    # subdivision_size 0 -> one histogram
    hn, sb, h = oracle.grsd21_subdiv(g, types, 0)
    assert hn == 1 and np.array_equal(h[0], h1)


@pytest.mark.parametrize("shape", ["plane", "sphere", "cylinder", "torus", "cone"])
def test_shape_data_known_answers(oracle, kat, shape):
    """Voxel occupancy + neighbour lookup against the counts decoded from the reference's
    *_GRSD_CCHLAC.pcd goldens (older revision: 13 half offsets, leaf 0.01)."""
    xyz = kat[f"{shape}_xyz"]
    counts = kat[f"{shape}_counts20"]
    nz = counts[counts != 0]
    g = oracle.voxel_grid(xyz, 0.01)
    g["leaf"] = 0.01
    V = g["nvox"]
    types = np.ones(V, np.int32)
    M, _ = oracle.grsd_transitions(g, types)
    pairs_half = M[1, 1] // 2
    empties_half = 13 * V - pairs_half
    assert M[1, 1] % 2 == 0 and M[1, 5] == 26 * V - 2 * pairs_half
    if shape == "cone":  # two classes in the golden; cross-class pairs were dropped by the old packing
        assert empties_half == nz[0] + nz[2] == 2113 and V == 308
    else:
        assert nz.tolist() == [empties_half, pairs_half]
        assert 13 * V == nz.sum()
    expected_V = {"plane": 325, "sphere": 391, "cylinder": 377, "torus": 147, "cone": 308}[shape]
    assert V == expected_V


def test_grsd21_pipeline_smoke(oracle, kat):
    out = oracle.grsd21(kat["sphere_xyz"], 0.01)
    assert out["nvox"] == 391 and out["hist21"].sum() == 391 * 26 and out["hist21"][20] == 0
    assert np.all((out["labels"] >= 0) & (out["labels"] <= 4))
    plane = oracle.grsd21(kat["plane_xyz"], 0.01)
    assert np.all(plane["labels"] == 1)  # exact plane -> every voxel PLANE
    assert plane["hist21"][6] == 2 * 1188 and plane["hist21"][10] == 26 * 325 - 2 * 1188


def test_grsd21_cluster_generator(oracle):
    xyz, off = synth.clusters(3, 1300, 2500)
    assert off[0] == 0 and off[-1] == xyz.shape[0]
    for c in range(3):
        out = oracle.grsd21(xyz[off[c]:off[c + 1]], 0.025)
        assert 0 < out["hist21"].sum() <= out["nvox"] * 26  # lower triangle is dropped


def _py_signature(kind, grid, types, cn, sub=0, off=(0, 0, 0)):
    """Independent pure-Python restatement of grsd_colorCHLAC_tools.hpp:305-451 / :462-668 (and :131-294)
    working on integer voxel coordinates recovered from the dense layout."""
    div = grid["div_b"].astype(int)
    lay = grid["layout"].reshape(div[2], div[1], div[0])
    coord = {int(lay[k, j, i]): (i, j, k) for k in range(div[2]) for j in range(div[1]) for i in range(div[0]) if lay[k, j, i] >= 0}
    offs = [(i, j, -1) for i in (-1, 0, 1) for j in (-1, 0, 1)] + [(i, -1, 0) for i in (-1, 0, 1)] + [(-1, 0, 0)]
    offs26 = offs + [(-a, -b, -c) for a, b, c in offs]
    if sub > 0:
        if any(div[a] <= off[a] for a in range(3)):
            return 0, None
        sb = [int(np.ceil((div[a] - off[a]) * np.float32(1.0 / sub))) for a in range(3)]
    else:
        sb = [1, 1, 1]
    hn = sb[0] * sb[1] * sb[2]
    dim = {0: 21, 1: 325, 2: 110}[kind]
    H = np.zeros((hn, dim), np.int64)
    tri6 = {(i, j): n for n, (i, j) in enumerate((i, j) for i in range(6) for j in range(i, 6))}
    tri5 = {(i, j): n for n, (i, j) in enumerate((i, j) for i in range(5) for j in range(i, 5))}
    unit = None
    if kind == 2:
        c32 = cn.astype(np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            ln = np.sqrt((c32[:, 0] * c32[:, 0] + c32[:, 1] * c32[:, 1]) + c32[:, 2] * c32[:, 2], dtype=np.float32)
            unit = c32 / ln[:, None]

    def nbr(v, o):
        i, j, k = coord[v]
        a, b, c = i + o[0], j + o[1], k + o[2]
        if 0 <= a < div[0] and 0 <= b < div[1] and 0 <= c < div[2]:
            return int(lay[c, b, a])
        return -1

    for v in range(grid["nvox"]):
        h = 0
        if hn != 1:
            t = [coord[v][a] - off[a] for a in range(3)]
            if min(t) < 0:
                continue
            q = [int(np.floor(np.float32(t[a]) * np.float32(1.0 / sub))) for a in range(3)]
            h = q[0] + q[1] * sb[0] + q[2] * sb[0] * sb[1]
        s = int(types[v])
        if kind == 0:
            for o in offs26:
                n = nbr(v, o)
                t = 5 if n < 0 else int(types[n])
                if s <= t:
                    H[h, tri6[(s, t)]] += 1
        elif kind == 1:
            for oid, o in enumerate(offs):
                n = nbr(v, o)
                if n >= 0:
                    H[h, s + 5 * int(types[n]) + 25 * oid] += 1
        else:
            if not np.all(np.isfinite(unit[v])):
                continue
            for o in offs26:
                n = nbr(v, o)
                if n < 0 or not np.all(np.isfinite(unit[n])):
                    H[h, 105 + s] += 1
                    continue
                a, b = unit[v], unit[n]
                cr = np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], np.float32)
                nm = np.sqrt((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2], dtype=np.float32)
                d = min(6, int(np.floor(np.sqrt(np.float64(nm)) * 7)))
                t = int(types[n])
                if s <= t:
                    H[h, d * 15 + tri5[(s, t)]] += 1
    return hn, H


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("sub,off", [(0, (0, 0, 0)), (2, (0, 0, 0)), (3, (1, 0, 2)), (2, (9, 0, 0))])
def test_signatures_match_python_restatement(oracle, kind, sub, off):
    rng = np.random.default_rng(100 + kind)
    leaf = 0.01
    occ = rng.random((6, 5, 4)) < 0.6
    ijk = np.argwhere(occ)
    reps = rng.integers(1, 4, size=len(ijk))
    base = np.repeat(ijk, reps, axis=0).astype(np.float64)
    pts = synth.quantize((base + rng.uniform(0.1, 0.9, size=base.shape)) * leaf + np.array([0.3, -0.2, 0.7])).astype(np.float32)
    nrm = rng.normal(size=pts.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[rng.integers(0, len(nrm), 5)] = np.nan  # voxels with a NaN mean normal
    g = oracle.voxel_grid(pts, leaf)
    g["leaf"] = leaf
    assert g["nvox"] == len(ijk)
    types = rng.integers(0, 5, size=g["nvox"]).astype(np.int32)
    cn = oracle.voxel_normals(pts, nrm, leaf)
    assert cn.shape == (g["nvox"], 3) and np.isnan(cn).any()
    hn, sb, h = oracle.grsd_signature(kind, g, types, cn, sub, off)
    phn, ph = _py_signature(kind, g, types, cn, sub, off)
    assert hn == phn
    if hn:
        assert np.array_equal(h, ph)
        assert h.sum() > 0
    if kind == 0:  # the general entry point agrees with the GRSD-21 one
        hn2, sb2, h2 = oracle.grsd21_subdiv(g, types, sub, off)
        assert hn2 == hn and np.array_equal(sb2, sb) and np.array_equal(h2, h)


def test_signature_hand_block(oracle):
    """3x3x3 full block, centre SPHERE(3), rest PLANE(1), all voxel normals +z except the centre (+x)."""
    leaf = 0.01
    ii, jj, kk = np.meshgrid(range(3), range(3), range(3), indexing="ij")
    pts = ((np.stack([ii, jj, kk], -1).reshape(-1, 3) + 0.5) * leaf).astype(np.float32)
    g = oracle.voxel_grid(pts, leaf)
    g["leaf"] = leaf
    types = np.ones(27, np.int32)
    centre = int(g["layout"][1 + 3 + 9])
    types[centre] = 3
    cn = np.tile(np.array([[0, 0, 2.0]], np.float32), (27, 1))  # un-normalised on purpose
    cn[centre] = [0.5, 0, 0]
    # GRSD-325: 13 half offsets; the centre sees all 13 (src 3 -> nbr 1); its 13 "upper" neighbours see it
    hn, _, h = oracle.grsd_signature(oracle.SIG_GRSD325, g, types, None)
    assert hn == 1 and h.shape == (1, 325)
    h = h[0].reshape(13, 5, 5)  # [offset, nbr, src]
    assert np.all(h[:, 1, 3] == 1) and np.all(h[:, 3, 1] == 1)
    occupied_pairs = 8 * 7 + 12 * 11 + 6 * 17 + 26
    assert h.sum() == occupied_pairs // 2 and h[:, 1, 1].sum() == occupied_pairs // 2 - 26
    # PlusGRSD-110: parallel normals -> angle bin 0; centre <-> others are orthogonal -> sin = 1 -> bin 6
    hn, _, p = oracle.grsd_signature(oracle.SIG_PLUSGRSD110, g, types, cn)
    p = p[0]
    tri5 = {(i, j): n for n, (i, j) in enumerate((i, j) for i in range(5) for j in range(i, 5))}
    assert p[0 * 15 + tri5[(1, 1)]] == occupied_pairs - 52
    assert p[6 * 15 + tri5[(1, 3)]] == 26  # (1 -> 3) ordered pairs; (3 -> 1) is below the diagonal and dropped
    assert p[105 + 1] == 27 * 26 - occupied_pairs and p[105 + 3] == 0
    assert p.sum() == (occupied_pairs - 52) + 26 + (27 * 26 - occupied_pairs)
    # a NaN source normal drops the voxel, a NaN neighbour normal counts as "to empty"
    cn2 = cn.copy()
    cn2[centre] = np.nan
    _, _, q = oracle.grsd_signature(oracle.SIG_PLUSGRSD110, g, types, cn2)
    q = q[0]
    assert q[6 * 15 + tri5[(1, 3)]] == 0 and q[105 + 1] == p[105 + 1] + 26 and q[105 + 3] == 0


def test_grsd_cluster_recipe_all_kinds(oracle):
    xyz, off = synth.clusters(2, 1300, 2000, seed_extra=5)
    pts = xyz[off[0]:off[1]]
    r21 = oracle.grsd_cluster(pts, 0.025, oracle.SIG_GRSD21)
    assert np.array_equal(r21["hist"][0], oracle.grsd21(pts, 0.025)["hist21"])
    r325 = oracle.grsd_cluster(pts, 0.025, oracle.SIG_GRSD325)
    r110 = oracle.grsd_cluster(pts, 0.025, oracle.SIG_PLUSGRSD110, subdivision_size=2)
    # every occupied (src, nbr) pair is seen once over the half offsets and twice over all 26
    occupied = r21["hist"][0].sum() - r21["hist"][0][[5, 10, 14, 17, 19]].sum()
    assert r325["hist"].sum() * 2 >= occupied  # 21 drops the below-diagonal ordered pairs
    assert r110["hist_num"] == int(np.prod(r110["subdiv_b"])) and r110["hist"].shape[1] == 110


def test_normals_against_the_references_own_output(oracle):
    """color_chlac/demos/data/tmp_normal.pcd carries the normals and curvatures the reference's computeNormal
    (pcl::NormalEstimation, radius 0.02, viewpoint 0) wrote for its 4712 points (mean 273 neighbours each): the oracle
    reproduces them -- direction AND sign -- which pins the neighbour rule, the PCA, curvature = l0 / (l0 + l1 + l2) and
    the flip towards the viewpoint against the reference itself."""
    g = np.load(GOLDEN / "tmp_normal.npz")
    xyz, ref_n, ref_c = g["xyz"], g["normal"], g["curvature"]
    n4, k = oracle.normals(xyz, 0.02)
    assert k.min() >= 3 and 250 < k.mean() < 300
    sin_angle = np.linalg.norm(np.cross(n4[:, :3].astype(np.float64), ref_n.astype(np.float64)), axis=1)
    assert (np.sum(n4[:, :3].astype(np.float64) * ref_n, axis=1) > 0).all()  # flipNormalTowardsViewpoint agrees everywhere
    assert np.sum(sin_angle > 1e-4) <= 1 and sin_angle.max() < 2e-4  # one point of curvature 0.18 (lambda0 close to lambda1)
    assert np.percentile(sin_angle, 99.9) < 3e-5
    assert np.abs(n4[:, 3] - ref_c).max() < 1e-6
    # and it is this recipe, not a neighbouring one
    for other in (oracle.normals(xyz, 0.03)[0], oracle.normals_knn(xyz, 30)):
        assert np.median(np.linalg.norm(np.cross(other[:, :3].astype(np.float64), ref_n.astype(np.float64)), axis=1)) > 1e-2
