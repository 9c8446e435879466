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
