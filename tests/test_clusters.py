"""Euclidean clustering (the segmentation step that produces GRSD's clusters):
the oracle against scipy's connected components and a pure-Python restatement of the reference's seed
loop, the GPU path against the oracle, and the chain clusters -> cab_grsd_batch."""
import numpy as np
import pytest

from mapping_private_b200 import cab, synth


def _objects_above_table(n=40_000, seed_sigma=0.0003):
    """The tabletop scene without its table: what table_object_detector_passive.cpp:270-287 hands to the clustering."""
    pts = synth.tabletop(n, noise_sigma=seed_sigma)
    return np.ascontiguousarray(pts[pts[:, 2] > 0.75 + 0.004])


def _scipy_labels(pts, tol, min_pts, max_pts=0):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree

    n = len(pts)
    fin = np.isfinite(pts).all(1)
    p64 = np.where(fin[:, None], pts, 1e9 + np.arange(n)[:, None] * 10.0).astype(np.float64)
    pairs = cKDTree(p64).query_pairs(tol, output_type="ndarray")
    g = coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    _, comp = connected_components(g, directed=False)
    size = np.bincount(comp)
    keep = fin & (size[comp] >= min_pts) & ((max_pts <= 0) | (size[comp] <= max_pts))
    out = np.full(n, -1, np.int32)
    ids = {}
    for i in range(n):  # clusters numbered by their smallest index
        if keep[i]:
            out[i] = ids.setdefault(comp[i], len(ids))
    return out, len(ids)


def _seed_loop(pts, tol, min_pts):
    """The published algorithm of extractEuclideanClusters, written naively (brute-force radius search)."""
    p = pts.astype(np.float32)
    n = len(p)
    r2 = np.float32(tol) * np.float32(tol)
    processed = np.zeros(n, bool)
    clusters = []
    for i in range(n):
        if processed[i] or not np.isfinite(p[i]).all():
            continue
        queue = [i]
        processed[i] = True
        sq = 0
        while sq < len(queue):
            d = p - p[queue[sq]]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            for j in np.nonzero((d2 <= r2) & ~processed)[0]:
                processed[j] = True
                queue.append(int(j))
            sq += 1
        if len(queue) >= min_pts:
            clusters.append(sorted(queue))
    return clusters


def test_oracle_clusters_against_scipy_and_seed_loop(oracle):
    pts = _objects_above_table()
    assert len(pts) > 10_000
    for tol, min_pts in [(0.02, 30), (0.05, 30), (0.004, 5)]:
        lab, nc = oracle.euclidean_clusters(pts, tol, min_pts)
        want, wnc = _scipy_labels(pts, tol, min_pts)
        assert nc == wnc and np.array_equal(lab, want), (tol, min_pts)
    lab, nc = oracle.euclidean_clusters(pts, 0.02, 30)
    assert nc == 5  # the five objects of the synthetic tabletop
    small = np.ascontiguousarray(pts[::12])
    small[3] = np.nan
    lab, nc = oracle.euclidean_clusters(small, 0.02, 4)
    clusters = _seed_loop(small, 0.02, 4)
    assert nc == len(clusters) and lab[3] == -1
    for c, members in enumerate(clusters):
        assert np.array_equal(np.nonzero(lab == c)[0], members)
    # max_pts drops the big ones
    lab2, nc2 = oracle.euclidean_clusters(pts, 0.02, 30, max_pts=int(np.bincount(lab[lab >= 0]).max()) - 1 if nc else 0)
    assert nc2 <= 5


def test_cluster_csr_is_host_only():
    """cab_cluster_csr needs no device: labels -> the reference's vector<vector<int>> layout."""
    L = cab.lib()
    import ctypes as C

    labels = np.array([1, -1, 0, 1, 0, 0, -1, 2], np.int32)
    off = np.zeros(4, np.int32)
    idx = np.zeros(6, np.int32)
    L.cab_cluster_csr.restype = C.c_int64
    total = L.cab_cluster_csr(labels.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(8), C.c_int32(3),
                              off.ctypes.data_as(C.POINTER(C.c_int32)), idx.ctypes.data_as(C.POINTER(C.c_int32)))
    assert total == 6 and off.tolist() == [0, 3, 5, 6] and idx.tolist() == [2, 4, 5, 0, 3, 7]
    assert L.cab_cluster_csr(labels.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(8), C.c_int32(2),
                             off.ctypes.data_as(C.POINTER(C.c_int32)), None) < 0  # a label beyond n_clusters


@pytest.mark.gpu
@pytest.mark.parametrize("tol,min_pts,max_pts", [(0.02, 30, 0), (0.05, 30, 0), (0.004, 5, 0), (0.003, 1, 0), (0.02, 30, 6000)])
def test_gpu_clusters_match_oracle(oracle, tol, min_pts, max_pts):
    ctx = cab.Context(0)
    pts = _objects_above_table()
    pts[11] = np.nan
    pts[500] = pts[501]  # a duplicate
    ctx.upload(pts)
    lab, nc = ctx.euclidean_clusters(tol, min_pts, max_pts)
    want, wnc = oracle.euclidean_clusters(pts, tol, min_pts, max_pts)
    assert nc == wnc
    assert np.array_equal(lab, want)
    assert lab[11] == -1
    off, idx = ctx.cluster_csr(lab, nc)
    assert off[-1] == (lab >= 0).sum()
    for c in range(nc):
        assert np.array_equal(idx[off[c]:off[c + 1]], np.nonzero(lab == c)[0])
    assert ctx.profile()["cluster_ms"] > 0


@pytest.mark.gpu
def test_gpu_clusters_whole_scene_and_edge_cases(oracle):
    ctx = cab.Context(0)
    pts = synth.tabletop(200_000, noise_sigma=0.0002)  # table and objects touch: one component
    ctx.upload(pts)
    lab, nc = ctx.euclidean_clusters(0.01, 10)
    want, wnc = oracle.euclidean_clusters(pts, 0.01, 10)
    assert nc == wnc and np.array_equal(lab, want)
    # sharded contexts cluster the whole cloud as well
    ctx.set_shard(1, 4)
    try:
        lab2, nc2 = ctx.euclidean_clusters(0.01, 10)
    finally:
        ctx.set_shard(0, 1)
    assert nc2 == nc and np.array_equal(lab2, lab)
    # isolated points, everything dropped, a single point, no finite point
    lone = synth.quantize(np.random.default_rng(3).uniform(-1, 1, (500, 3))).astype(np.float32)
    ctx.upload(lone)
    lab, nc = ctx.euclidean_clusters(1e-4, 1)
    assert nc == 500 and np.array_equal(lab, np.arange(500))
    lab, nc = ctx.euclidean_clusters(1e-4, 2)
    assert nc == 0 and (lab == -1).all()
    ctx.upload(lone[:1])
    lab, nc = ctx.euclidean_clusters(0.05, 1)
    assert nc == 1 and lab.tolist() == [0]
    ctx.upload(np.full((4, 3), np.nan, np.float32))
    lab, nc = ctx.euclidean_clusters(0.05, 1)
    assert nc == 0 and (lab == -1).all()
    with pytest.raises(cab.CabError):
        ctx.euclidean_clusters(0.0, 1)


@pytest.mark.gpu
def test_clusters_feed_grsd_batch(oracle):
    """The reference's chain: objects above the table -> Euclidean clusters -> one GRSD histogram per cluster."""
    ctx = cab.Context(0, exact=True)
    pts = _objects_above_table(60_000, seed_sigma=0.0)
    ctx.upload(pts)
    lab, nc = ctx.euclidean_clusters(0.02, 30)
    assert nc == 5
    off, idx = ctx.cluster_csr(lab, nc)
    xyz = np.ascontiguousarray(pts[idx])
    hist = ctx.grsd_batch(xyz, off, 0.025)
    for c in range(nc):
        o = oracle.grsd21(xyz[off[c]:off[c + 1]], 0.025)
        assert np.array_equal(hist[c], o["hist21"])


def test_host_mirror_rejects_the_normal_angle_variant():
    """The free function keeps the reference signature; the region-growing variant (nx_idx >= 0) is reported, not guessed."""
    from mapping_private_b200 import plugin

    plugin.build()
    pts = synth.analytic_shape("plane", 200)
    assert plugin.extract_euclidean_clusters(pts, np.arange(200), 0.05, 10, nx_idx=0) is None
    assert plugin.extract_euclidean_clusters(pts, np.zeros(0, np.int32), 0.05, 10) == []


@pytest.mark.gpu
def test_host_mirror_clusters_a_subset_like_the_reference_call(oracle):
    """extractEuclideanClusters(points, object_indices, 0.05, clusters, -1, -1, -1, -1, 30) as at
    table_object_detector_passive.cpp:293: only the points named by `indices` take part, clusters hold cloud indices."""
    from mapping_private_b200 import plugin

    plugin.build()
    pts = synth.tabletop(40_000, noise_sigma=0.0003)
    indices = np.nonzero(pts[:, 2] > 0.754)[0].astype(np.int32)
    clusters = plugin.extract_euclidean_clusters(pts, indices, 0.02, 30)
    lab, nc = oracle.euclidean_clusters(pts[indices], 0.02, 30)
    assert clusters is not None and len(clusters) == nc == 5
    for c in range(nc):
        assert np.array_equal(clusters[c], indices[lab == c])
