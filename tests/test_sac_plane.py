"""Table-plane step: fitSACPlane = MSAC over a plane model + least-squares refinement + inlier selection +
projection (cloud_tools/src/table_object_detector_passive.cpp:621-659; sample_consensus is external: the published
algorithm on a GIVEN sample sequence).  CPU: the oracle against analytic planes.  GPU: cab_fit_plane_msac against the
oracle on the same sequence -- same winning hypothesis, same iteration count, same inliers."""
import numpy as np
import pytest

import pyoracle
from mapping_private_b200 import cab, synth


def _scene(n_plane=20_000, n_clutter=8_000, seed=3, tilt=(0.05, -0.03), z0=0.72, noise=0.002):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-0.6, 0.6, (n_plane, 2))
    z = z0 + tilt[0] * xy[:, 0] + tilt[1] * xy[:, 1] + rng.normal(0, noise, n_plane)
    plane = np.column_stack([xy, z])
    clutter = np.column_stack([rng.uniform(-0.6, 0.6, (n_clutter, 2)), rng.uniform(z0 + 0.05, z0 + 0.45, n_clutter)])
    pts = np.concatenate([plane, clutter]).astype(np.float32)
    return pts[rng.permutation(pts.shape[0])], np.array([-tilt[0], -tilt[1], 1.0, -z0]) / np.sqrt(1 + tilt[0] ** 2 + tilt[1] ** 2)


def _triples(m, count, seed=11):
    rng = np.random.default_rng(seed)
    return np.stack([rng.choice(m, 3, replace=False) for _ in range(count)]).astype(np.int32)


def test_oracle_recovers_the_plane():
    pts, truth = _scene()
    tri = _triples(pts.shape[0], 501)
    r = pyoracle.fit_plane_msac(pts, tri, threshold=0.01)
    c = r["coeff"] * np.sign(r["coeff"][2])
    assert np.allclose(c, truth, atol=2e-3)
    assert 19_000 < r["inliers"].shape[0] < 21_500
    assert 0 <= r["best_iteration"] < 501 and 0 < r["iterations"] <= 501
    assert r["iterations"] < 60  # w ~ 0.7: k = log(0.01) / log(1 - 0.7^3) ~ 11 once a good model is found
    d = np.abs(pts[r["inliers"]].astype(np.float64) @ r["coeff"][:3] + r["coeff"][3])
    assert d.max() <= 0.01
    assert np.abs(r["projected"].astype(np.float64) @ r["coeff"][:3] + r["coeff"][3]).max() < 1e-6


def test_oracle_degenerate_cases():
    pts, _ = _scene(500, 100)
    # repeated and collinear samples are iterations without a model; fewer than three points: no model
    tri = np.array([[0, 0, 1], [1, 2, 2]] + _triples(pts.shape[0], 50).tolist(), np.int32)
    r = pyoracle.fit_plane_msac(pts, tri, threshold=0.01)
    assert r["best_iteration"] >= 2
    r2 = pyoracle.fit_plane_msac(pts[:2], np.array([[0, 1, 1]], np.int32))
    assert r2["inliers"].shape[0] == 0 and r2["best_iteration"] == -1 and np.all(r2["coeff"] == 0)
    line = np.column_stack([np.linspace(0, 1, 50), np.zeros(50), np.zeros(50)]).astype(np.float32)
    r3 = pyoracle.fit_plane_msac(line, _triples(50, 20))
    assert r3["best_iteration"] == -1 and r3["iterations"] == 0  # no valid hypothesis: nothing was scored


def _agree(g, o, pts):
    assert g["best_iteration"] == o["best_iteration"] and g["iterations"] == o["iterations"]
    assert np.array_equal(g["inliers"], o["inliers"])
    assert np.allclose(g["coeff"], o["coeff"], rtol=0, atol=1e-10)
    assert np.abs(g["projected"] - o["projected"]).max() <= 1e-6 if o["projected"].size else True


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["all", "subset", "big", "few_samples", "degenerate_first"])
def test_gpu_equals_oracle(case):
    ctx = cab.Context(0)
    if case == "big":
        pts, _ = _scene(300_000, 120_000, seed=5)
    else:
        pts, _ = _scene()
    idx = None
    m = pts.shape[0]
    if case == "subset":  # the detector fits one cluster of a larger cloud
        idx = np.flatnonzero(pts[:, 2] < 0.9).astype(np.int32)[::2]
        m = idx.shape[0]
    tri = _triples(m, 8 if case == "few_samples" else 501)
    if case == "degenerate_first":
        tri[0] = [5, 5, 9]
        tri[1] = [7, 3, 3]
    thr = 0.03 if case == "big" else 0.01
    o = pyoracle.fit_plane_msac(pts, tri, indices=idx, threshold=thr)
    g = ctx.fit_plane_msac(pts, tri, indices=idx, threshold=thr)
    _agree(g, o, pts)
    assert g["inliers"].shape[0] > 0.5 * (m if case != "subset" else 1)
    # too few points / bad arguments
    e = ctx.fit_plane_msac(pts[:2], np.array([[0, 1, 1]], np.int32))
    assert e["inliers"].shape[0] == 0 and e["best_iteration"] == -1
    with pytest.raises(cab.CabError):
        ctx.fit_plane_msac(pts, np.array([[0, 1, m + 5]], np.int32), indices=idx)
    ctx.close()


@pytest.mark.gpu
def test_host_function_projects_in_place():
    """cloud_tools::fitSACPlane with the member's signature (host/include/cloud_tools/fit_sac_plane.h)."""
    from mapping_private_b200 import plugin

    pts, truth = _scene(15_000, 4_000, seed=9)
    idx = np.arange(pts.shape[0], dtype=np.int32)
    rc, inl, coeff, out = plugin.fit_sac_plane(pts, idx, threshold=0.01, min_pts=10, seed=4)
    assert rc == inl.shape[0] > 14_000
    c = coeff * np.sign(coeff[2])
    assert np.allclose(c, truth, atol=2e-3)
    assert np.abs(out[inl].astype(np.float64) @ coeff[:3] + coeff[3]).max() < 1e-6  # projected in place
    rest = np.setdiff1d(idx, inl)
    assert np.array_equal(out[rest], pts[rest])
    rc2, _, _, _ = plugin.fit_sac_plane(pts, idx[:5], min_pts=10)
    assert rc2 == -1
