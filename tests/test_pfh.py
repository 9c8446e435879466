"""PointFeatureHistogram: the oracle's pair features against an independent numpy
restatement of cloud_algos/include/cloud_algos/pfh.h:102-238 and analytic cases, its histograms against their
invariants, the GPU path against the oracle."""
import numpy as np
import pytest

from mapping_private_b200 import cab, synth


def _pair_numpy(ps, ns, pt, nt, check_flip=True, abs_angles=False, max_dist=1.0):
    ps, ns, pt, nt = (np.asarray(v, np.float32) for v in (ps, ns, pt, nt))
    d = (pt - ps).astype(np.float64)
    delta = np.linalg.norm(d)
    if delta == 0:
        return None
    angle2 = -np.dot(nt.astype(np.float64), d) / delta
    src, tgt = ns.astype(np.float64), nt.astype(np.float64)
    gamma = np.dot(src, d) / delta
    flip = (not check_flip) or (np.arccos(np.clip(gamma, -1, 1)) > np.arccos(np.clip(angle2, -1, 1)))
    if flip:
        src, tgt, d, gamma = tgt, src, -d, angle2
    u = src
    tmp = np.cross(d, u)
    nrm = np.linalg.norm(tmp)
    if nrm == 0:
        return None
    v = tmp / nrm
    w = np.cross(u, v)
    beta = np.dot(v, tgt)
    if abs_angles:
        return np.array([np.arctan2(abs(np.dot(w, tgt)), abs(np.dot(u, tgt))) / (np.pi / 2), abs(beta), abs(gamma), delta / max_dist])
    alpha = np.arctan2(np.dot(w, tgt), np.dot(u, tgt))
    return np.array([(alpha + np.pi) / (2 * np.pi), (beta + 1) / 2, (gamma + 1) / 2, delta / max_dist])


def test_pair_features_match_numpy_restatement(oracle):
    rng = np.random.default_rng(0)
    for t in range(400):
        ps, pt = rng.normal(size=3) * 0.02, rng.normal(size=3) * 0.02
        ns, nt = rng.normal(size=3), rng.normal(size=3)
        ns, nt = ns / np.linalg.norm(ns), nt / np.linalg.norm(nt)
        ps32, pt32 = ps.astype(np.float32), pt.astype(np.float32)
        dd = pt32 - ps32
        d2 = np.float32(np.float32(dd[0] * dd[0] + dd[1] * dd[1]) + dd[2] * dd[2])
        for cf, ab in ((True, False), (False, False), (True, True)):
            ok, f = oracle.pfh_pair(ps32, ns, pt32, nt, d2, 0.06, cf, ab)
            want = _pair_numpy(ps32, ns, pt32, nt, cf, ab, 0.06)
            assert ok and np.allclose(f, want, rtol=1e-6, atol=2e-7), (t, cf, ab, f, want)  # delta is an fp32 sqrt of the fp32 d2
    # invalid pairs: coincident points, normal parallel to the connecting line (pfh.h:117-121,171-175)
    assert not oracle.pfh_pair([0, 0, 0], [0, 0, 1], [0, 0, 0], [0, 0, 1], 0.0, 1.0)[0]
    assert not oracle.pfh_pair([0, 0, 0], [0, 0, 1], [0, 0, 0.01], [0, 0, 1], 1e-4, 1.0, check_flip=False)[0]
    # two points of a plane with parallel normals: alpha = 0 -> 0.5, beta = 0 -> 0.5, gamma = 0 -> 0.5
    ok, f = oracle.pfh_pair([0, 0, 0], [0, 0, 1], [0.01, 0, 0], [0, 0, 1], 1e-4, 0.06)
    assert ok and np.allclose(f[:3], [0.5, 0.5, 0.5]) and abs(f[3] - 0.01 / 0.06) < 1e-6


def test_histogram_invariants(oracle):
    pts = synth.tabletop(8000, noise_sigma=0.0003)
    nrm = np.nan_to_num(oracle.normals(pts, 0.02)[0][:, :3], nan=0.0)
    off, _, _ = oracle.radius_search(pts, pts, 0.03, max_nn=100)
    k = np.diff(off)
    spfh = oracle.pfh(pts, nrm, flags=oracle.PFH_CHECK_FLIP)
    # every feature's histogram of a point sums to 100 (k - 1) / k (pfh.cpp:212,267-271)
    for ft in range(3):
        assert np.allclose(spfh[:, 9 * ft:9 * ft + 9].sum(1), 100.0 * (k - 1) / k, rtol=1e-5)
    with_dist = oracle.pfh(pts, nrm, flags=oracle.PFH_CHECK_FLIP | oracle.PFH_USE_DIST)
    assert with_dist.shape[1] == 36 and np.array_equal(with_dist[:, :27], spfh)
    # delta / (2 r) <= 0.5: only the lower bins, except where an invalid pair (a zeroed normal) spreads its increment
    assert np.mean(with_dist[:, 27 + 5:].sum(1) > 1e-6) < 0.01
    # differential: cumulative sums give the plain histograms back
    diff = oracle.pfh(pts, nrm, flags=oracle.PFH_CHECK_FLIP | oracle.PFH_DIFFERENTIAL)
    for ft in range(3):
        assert np.allclose(np.cumsum(diff[:, 9 * ft:9 * ft + 9], axis=1), spfh[:, 9 * ft:9 * ft + 9], atol=1e-3)
    # FPFH: a convex combination of the neighbours' histograms keeps the per-feature mass near 100
    fpfh = oracle.pfh(pts, nrm)
    m = fpfh[:, :9].sum(1)
    assert np.all((m > 85) & (m < 100.001))


def test_combined_histogram_invariants(oracle):
    """combine_ = true (pfh.cpp:47-57, 239-258): one n-D histogram; its marginals are the 1-D histograms."""
    pts = synth.tabletop(4000, noise_sigma=0.0003)
    nrm = np.nan_to_num(oracle.normals(pts, 0.02)[0][:, :3], nan=0.0)
    q = 5
    sep = oracle.pfh(pts, nrm, quantum=q, flags=oracle.PFH_CHECK_FLIP)
    com = oracle.pfh(pts, nrm, quantum=q, flags=oracle.PFH_CHECK_FLIP | oracle.PFH_COMBINE)
    assert com.shape == (pts.shape[0], q ** 3)
    cube = com.reshape(-1, q, q, q)  # index = fi[0] + q fi[1] + q^2 fi[2]: axes (fi[2], fi[1], fi[0]) = (alpha, gamma, beta)
    assert np.allclose(cube.sum((2, 3)), sep[:, 0:q], atol=2e-3)        # alpha
    assert np.allclose(cube.sum((1, 2)), sep[:, q:2 * q], atol=2e-3)    # beta
    assert np.allclose(cube.sum((1, 3)), sep[:, 2 * q:3 * q], atol=2e-3)  # gamma
    com4 = oracle.pfh(pts, nrm, quantum=3, flags=oracle.PFH_CHECK_FLIP | oracle.PFH_COMBINE | oracle.PFH_USE_DIST)
    assert com4.shape[1] == 81
    # the differential option is ignored in the combined mode (:345)
    assert np.array_equal(com, oracle.pfh(pts, nrm, quantum=q, flags=oracle.PFH_CHECK_FLIP | oracle.PFH_COMBINE | oracle.PFH_DIFFERENTIAL))


@pytest.mark.gpu
@pytest.mark.parametrize("flags,max_nn,quantum", [(4 | 32, 100, 9), (4 | 16 | 32, 100, 5), (4 | 16 | 1 | 32, 0, 4), (8 | 32 | 2, 60, 6)])
def test_gpu_combined_pfh_matches_oracle(oracle, flags, max_nn, quantum):
    ctx = cab.Context(0)
    pts = synth.tabletop(9_000, noise_sigma=0.0003, seed_extra=flags)
    nrm = np.nan_to_num(oracle.normals(pts, 0.02)[0][:, :3], nan=0.0)
    ctx.upload(pts)
    ctx.build_grid(0.03)
    ctx.set_normals(nrm)
    got = ctx.pfh(0.03, max_nn, quantum, flags)
    want = oracle.pfh(pts, nrm, 0.03, max_nn, quantum, flags)
    assert got.shape == want.shape == (pts.shape[0], quantum ** (4 if flags & 1 else 3))
    if not flags & 16:
        assert np.mean(np.any(got != want, axis=1)) < 3e-3  # a feature within an ulp of a bin edge moves one count
        assert np.allclose(got.sum(1), want.sum(1), rtol=1e-5)
    else:
        assert np.allclose(got, want, rtol=2e-4, atol=2e-3)
    with pytest.raises(cab.CabError, match="at most 4096"):
        ctx.pfh(0.03, max_nn, 17, 4 | 32)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("flags,max_nn,quantum", [(4, 100, 9), (4 | 16, 100, 9), (4 | 16 | 1, 0, 7), (8 | 16 | 2, 60, 5), (16, 100, 9)])
def test_gpu_pfh_matches_oracle(oracle, flags, max_nn, quantum):
    ctx = cab.Context(0)
    pts = synth.tabletop(12_000, noise_sigma=0.0003, seed_extra=flags)
    nrm = np.nan_to_num(oracle.normals(pts, 0.02)[0][:, :3], nan=0.0)
    ctx.upload(pts)
    ctx.build_grid(0.03)
    ctx.set_normals(nrm)
    got = ctx.pfh(0.03, max_nn, quantum, flags)
    want = oracle.pfh(pts, nrm, 0.03, max_nn, quantum, flags)
    assert got.shape == want.shape
    if not flags & 16:
        # star histograms: integer counts times a constant increment, reproduced exactly unless a feature sits within
        # an ulp of a bin edge (device acos / atan2 differ from libm in the last bit)
        assert np.mean(np.any(got != want, axis=1)) < 1e-3
        assert np.allclose(got, want, atol=100.0 / 10 * 3)
    else:
        # the weighted average adds its terms in another order than the reference (ascending distance)
        assert np.allclose(got, want, rtol=2e-4, atol=2e-3)
    ctx.close()
