"""StatisticalNoiseRemoval's k-NN mean distances: the oracle against scipy's
k-d tree and a brute-force restatement, the GPU path against the oracle."""
import numpy as np
import pytest

from mapping_private_b200 import cab, synth


def _cloud(seed=1, n=20_000, outliers=200):
    rng = np.random.default_rng(seed)
    pts = synth.tabletop(n, noise_sigma=0.0005)
    far = synth.quantize(rng.uniform([-0.6, -0.4, 0.5], [0.6, 0.4, 1.3], (outliers, 3)))
    return np.concatenate([pts, far]).astype(np.float32)


def test_oracle_knn_against_brute_force_and_scipy(oracle):
    from scipy.spatial import cKDTree

    pts = _cloud(n=3000, outliers=60)
    pts[5] = pts[6]  # a duplicate: its nearest "other" neighbour is at distance 0
    k = 7
    avg = oracle.knn_mean_distance(pts, k)
    # brute force with the documented d2 rule, ties by index
    p = pts.astype(np.float32)
    ref = np.zeros(len(p))
    for i in range(0, len(p), 500):
        d = p[i:i + 500, None, :] - p[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        part = np.sort(d2, axis=1)[:, :k]
        ref[i:i + 500] = np.sqrt(part[:, 1:]).astype(np.float64).sum(1) / (k - 1)  # fp32 sqrt, fp64 sum
    assert np.allclose(avg, ref, rtol=1e-13, atol=0)
    dist, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=k)
    assert np.allclose(avg, dist[:, 1:].sum(1) / (k - 1), rtol=1e-5)
    keep, mean, std = oracle.noise_filter(avg, 3.0)
    assert abs(mean - avg.mean()) < 1e-15 and abs(std - avg.std()) < 1e-12
    assert np.array_equal(keep, np.abs(avg - mean) < 3.0 * std)
    # the plugin's argument checks (noise_removal.cpp:51-62)
    with pytest.raises(ValueError):
        oracle.knn_mean_distance(pts, 1)
    with pytest.raises(ValueError):
        oracle.knn_mean_distance(pts[:4], 10)
    # non-finite points: NaN, outside the statistics, never kept
    q = pts.copy()
    q[11] = np.nan
    a2 = oracle.knn_mean_distance(q, k)
    assert np.isnan(a2[11]) and not oracle.noise_filter(a2, 3.0)[0][11]


@pytest.mark.gpu
@pytest.mark.parametrize("k,hint", [(10, 0.0), (2, 0.0), (25, 0.004), (10, 0.5)])
def test_gpu_knn_mean_distance_matches_oracle(oracle, k, hint):
    ctx = cab.Context(0)
    pts = _cloud(seed=k)
    pts[7] = pts[8]
    pts[100] = np.inf
    ctx.upload(pts)
    avg = ctx.knn_mean_distance(k, cell_hint=hint)
    want = oracle.knn_mean_distance(pts, k)
    assert np.array_equal(np.isnan(avg), np.isnan(want)) and np.isnan(avg[100])
    fin = ~np.isnan(want)
    # same neighbours (bit-exact d2), fp64 sums in a different order
    assert np.max(np.abs(avg[fin] - want[fin]) / np.maximum(want[fin], 1e-30)) < 1e-13
    res = ctx.statistical_outliers(k, 3.0, cell_hint=hint)
    keep, mean, std = oracle.noise_filter(want, 3.0)
    assert abs(res["mean"] - mean) < 1e-15 and abs(res["stddev"] - std) < 1e-13
    sure = np.abs(np.abs(want - mean) - 3.0 * std) > 1e-10
    assert np.array_equal(res["keep"][sure & fin], keep[sure & fin]) and res["kept"] == int(res["keep"].sum())
    assert ctx.profile()["knn_rounds"] >= 1
    # argument checks mirror the plugin's
    with pytest.raises(cab.CabError):
        ctx.knn_mean_distance(1)
    ctx.upload(pts[:5])
    with pytest.raises(cab.CabError, match="nearest neighbors requested"):
        ctx.knn_mean_distance(10)
    ctx.close()


@pytest.mark.gpu
def test_gpu_knn_isolated_points_need_several_rounds(oracle):
    """Two far-apart blobs plus a lone point: the lone point's neighbours are metres away."""
    rng = np.random.default_rng(3)
    a = synth.quantize(rng.normal([0, 0, 0], 0.01, (3000, 3)))
    b = synth.quantize(rng.normal([4, 1, -2], 0.02, (3000, 3)))
    pts = np.concatenate([a, b, np.array([[2.0, 8.0, 3.0]], np.float32)]).astype(np.float32)
    ctx = cab.Context(0)
    ctx.upload(pts)
    avg = ctx.knn_mean_distance(10)
    want = oracle.knn_mean_distance(pts, 10)
    assert np.max(np.abs(avg - want) / want) < 1e-13
    assert ctx.profile()["knn_rounds"] > 3 and avg[-1] > 5.0
    ctx.close()


def _angle_to(a, b):
    """sin of the angle between two unit-vector arrays, sign-insensitive (acos of the dot product cannot resolve 1e-4 rad)."""
    return np.linalg.norm(np.cross(a.astype(np.float64), b.astype(np.float64)), axis=1)


def test_oracle_knn_normals_against_numpy(oracle):
    """k-NN normals of the oracle (table_object_detector_passive.cpp:668-714) against scipy's k-d tree + numpy PCA,
    and analytic planes."""
    from scipy.spatial import cKDTree

    pts = _cloud(n=4000, outliers=0)
    k = 10
    n4 = oracle.normals_knn(pts, k)
    p64 = pts.astype(np.float64)
    _, idx = cKDTree(p64).query(p64, k=k)
    ref = np.zeros((len(pts), 3))
    curv = np.zeros(len(pts))
    for i in range(len(pts)):
        nb = p64[idx[i]]
        w, v = np.linalg.eigh(np.cov(nb.T, bias=True))
        ref[i] = v[:, 0]
        curv[i] = w[0] / w.sum()
    good = curv < 0.05  # well-conditioned neighbourhoods (an edge point's smallest two eigenvalues can be close)
    assert good.mean() > 0.7
    assert np.percentile(_angle_to(n4[good, :3], ref[good]), 99) < 1e-3  # scipy breaks distance ties its own way
    assert np.allclose(n4[good, 3], curv[good], atol=1e-4)
    # flipped towards the viewpoint (0, 0, 0)
    assert (np.sum(n4[:, :3].astype(np.float64) * (-p64), axis=1) >= -1e-7).all()
    plane = synth.analytic_shape("plane", 2000)
    np4 = oracle.normals_knn(plane, 12, vp=(0.0, 0.0, 10.0))
    assert np.abs(np.abs(np4[:, 2]) - 1).max() < 1e-6 and np.abs(np4[:, 3]).max() < 1e-9
    with pytest.raises(ValueError):
        oracle.normals_knn(pts, 2)
    with pytest.raises(ValueError):
        oracle.normals_knn(pts[:5], 10)


@pytest.mark.gpu
@pytest.mark.parametrize("k,hint,exact", [(10, 0.0, False), (10, 0.0, True), (30, 0.003, False), (5, 0.4, True)])
def test_gpu_knn_normals_match_oracle(oracle, k, hint, exact):
    ctx = cab.Context(0, exact=exact)
    pts = _cloud(seed=k + 1)
    pts[7] = pts[8]
    pts[100] = np.inf
    ctx.upload(pts)
    vp = (0.1, -0.2, 2.0)
    n4 = ctx.normals_knn(k, vp=vp, cell_hint=hint)
    want = oracle.normals_knn(pts, k, vp=vp)
    assert np.isnan(n4[100]).all() and np.isnan(want[100]).all()
    fin = ~np.isnan(want[:, 0])
    assert np.array_equal(~np.isnan(n4[:, 0]), fin)
    ang = _angle_to(n4[fin, :3], want[fin, :3])
    well = want[fin, 3] < 0.1
    if exact:
        assert np.percentile(ang[well], 99.9) < 1e-4
    else:
        assert np.mean(ang[well] > 1e-4) < 5e-3  # fp32 sums over 5..30 points: the tail is ill-conditioned neighbourhoods
    assert np.allclose(n4[fin, 3][well], want[fin, 3][well], atol=2e-4)
    # the flip is decided by the same viewpoint
    d = (np.asarray(vp) - pts[fin].astype(np.float64))
    assert (np.sum(n4[fin, :3] * d, axis=1) >= -1e-6).all()
    assert ctx.profile()["knn_rounds"] >= 1


@pytest.mark.gpu
def test_gpu_knn_normals_feed_rsd_and_errors(oracle):
    ctx = cab.Context(0)
    pts = synth.tabletop(30_000, noise_sigma=0.0003)
    ctx.upload(pts)
    n4 = ctx.normals_knn(25)
    ctx.build_grid(0.02)
    ctx.set_normals(n4)
    rmin, rmax = ctx.rsd(0.02)
    omin, omax, _ = oracle.rsd(pts, n4, 0.02)
    assert np.max(np.abs(rmin - omin) / omin) < 1e-4 and np.max(np.abs(rmax - omax) / omax) < 1e-4
    with pytest.raises(cab.CabError):
        ctx.normals_knn(2)
    ctx.upload(pts[:6])
    with pytest.raises(cab.CabError):
        ctx.normals_knn(10)
