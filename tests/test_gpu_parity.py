"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on
the same seeded inputs.  Bit-exact for neighbour sets, counts, voxel keys, labels and histograms;
normals within 1e-4 rad (sign-insensitive); RSD radii within 1e-4 relative."""
import numpy as np
import pytest

from mapping_private_b200 import cab, synth

pytestmark = pytest.mark.gpu

NORMAL_TOL_RAD = 1e-4
RADIUS_TOL_REL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    c = cab.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_exact():
    c = cab.Context(0, exact=True)
    yield c
    c.close()


@pytest.fixture(scope="module")
def c1():
    return synth.tabletop(100_000)


def _canon(off, idx, d2):
    """Sort every query's list by (d2, idx)."""
    q = np.repeat(np.arange(len(off) - 1), np.diff(off))
    order = np.lexsort((idx, d2, q))
    return idx[order], d2[order]


def _angle(a, b):
    """sin of the angle between two unit-vector arrays, sign-insensitive."""
    return np.linalg.norm(np.cross(a.astype(np.float64), b.astype(np.float64)), axis=1)


@pytest.mark.parametrize("max_nn", [0, 75])
def test_neighbor_sets_bit_exact_c1(ctx, oracle, c1, max_nn):
    r = 0.02
    ctx.upload(c1)
    ctx.build_grid(r)
    n = c1.shape[0]
    q0, q1 = (0, n) if max_nn == 0 else (0, 30_000)
    off, idx, d2 = ctx.neighbors(r, q0, q1, max_nn=max_nn)
    ooff, oidx, od2 = oracle.radius_search(c1, c1[q0:q1], r, max_nn=max_nn)
    assert np.array_equal(off, ooff)
    gi, gd = _canon(off, idx, d2)
    assert np.array_equal(gi, oidx)
    assert np.array_equal(gd.view(np.uint32), od2.view(np.uint32))


def test_neighbor_sets_radius_boundary(ctx, oracle):
    # points exactly at distance r (d2 == r2 in fp32) are neighbours; one ulp beyond is not
    r = np.float32(0.0625)
    rng = np.random.default_rng(7)
    base = synth.quantize(rng.uniform(0.2, 0.4, size=(2000, 3)))
    extra = []
    for p in base[:200]:
        extra.append(p + np.array([r, 0, 0], np.float32))
        extra.append(p + np.array([0, np.nextafter(r, np.float32(1)), 0], np.float32))
        extra.append(p - np.array([0, 0, r], np.float32))
    pts = np.concatenate([base, np.array(extra, np.float32)]).astype(np.float32)
    ctx.upload(pts)
    ctx.build_grid(float(r))
    off, idx, d2 = ctx.neighbors(float(r), 0, pts.shape[0])
    ooff, oidx, od2 = oracle.radius_search(pts, pts, float(r))
    assert np.array_equal(off, ooff)
    gi, gd = _canon(off, idx, d2)
    assert np.array_equal(gi, oidx) and np.array_equal(gd, od2)
    assert np.any(gd == r * r)


@pytest.mark.parametrize("exact", [False, True])
def test_normals_parity_c1(ctx, ctx_exact, oracle, c1, exact):
    c = ctx_exact if exact else ctx
    r = 0.02
    c.upload(c1)
    c.build_grid(r)
    n4 = c.normals(r)
    o4, ok = oracle.normals(c1, r)
    prof = c.profile()
    assert prof["neighbour_sum"] == int(ok.sum())  # neighbour counts are bit-exact
    nan_g, nan_o = np.isnan(n4[:, 0]), np.isnan(o4[:, 0])
    assert np.array_equal(nan_g, nan_o)
    good = ~nan_o
    ang = _angle(n4[good, :3], o4[good, :3])
    tol = 2e-6 if exact else NORMAL_TOL_RAD
    bad = ang > tol
    # ill-conditioned neighbourhoods (two smallest eigenvalues nearly equal) are listed, not hidden
    print(f"normals exact={exact}: max sin(angle) {ang.max():.3e}, > tol: {bad.sum()} of {good.sum()}")
    assert bad.mean() < 1e-4
    assert np.allclose(np.linalg.norm(n4[good, :3], axis=1), 1.0, atol=1e-5)
    assert np.max(np.abs(n4[good, 3] - o4[good, 3])) < (1e-6 if exact else 2e-5)
    # orientation: towards the viewpoint unless the oracle's own decision margin is tiny
    dots = np.sum(n4[good, :3] * o4[good, :3], axis=1)
    margin = np.abs(np.sum(o4[good, :3] * (-c1[good]), axis=1))
    assert np.all((dots > 0) | (margin < 1e-4))


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("max_nn,ndiv,plane,flags", [(0, 10, 0.1, 0), (75, 10, 0.1, 0), (150, 10, 0.1, 0), (12, 10, 0.1, 2), (0, 5, 0.2, 2 | 4),
                                                     (0, 7, 0.15, 4)])
def test_rsd_parity_given_normals(ctx, ctx_exact, oracle, exact, max_nn, ndiv, plane, flags):
    c = ctx_exact if exact else ctx
    pts = synth.tabletop(40_000, noise_sigma=0.0004)
    r = 0.03
    o4, _ = oracle.normals(pts, r)
    c.upload(pts)
    c.build_grid(r)
    c.set_normals(o4)
    rmin, rmax = c.rsd(r, max_nn=max_nn, ndiv=ndiv, plane_radius=plane, flags=flags)
    omin, omax, _ = oracle.rsd(pts, o4, r, max_nn=max_nn, ndiv=ndiv, plane_radius=plane, flags=flags)
    tol = 1e-6 if exact else RADIUS_TOL_REL
    assert np.max(np.abs(rmin - omin) / omin) <= tol
    assert np.max(np.abs(rmax - omax) / omax) <= tol
    if exact:
        assert np.mean(rmin == omin) > 0.999 and np.mean(rmax == omax) > 0.999


def test_rsd_truncation_with_ties_and_missing_normals(ctx, oracle):
    """The max_nn paths of the fast mode: a lattice (every distance occurs many times: the target bin of the d2 histogram
    overflows its list and the packet takes the exact-threshold path) and a cloud with NaN normals (such neighbours count
    for the truncation but form no pair), against the oracle's (d2, index) rule."""
    g = np.arange(0, 0.12, 0.004, dtype=np.float32)
    lattice = np.stack(np.meshgrid(g, g, g[:6], indexing="ij"), axis=-1).reshape(-1, 3)
    noisy = synth.tabletop(20_000, noise_sigma=0.0004)
    for pts, max_nn in ((lattice, 40), (noisy, 60), (noisy, 200)):
        r = 0.02
        o4, _ = oracle.normals(pts, r)
        o4[::37, :3] = np.nan
        ctx.upload(pts)
        ctx.build_grid(r)
        ctx.set_normals(o4)
        rmin, rmax = ctx.rsd(r, max_nn=max_nn)
        k_gpu = ctx.profile()["neighbour_sum"]
        omin, omax, _ = oracle.rsd(pts, o4, r, max_nn=max_nn)
        assert np.max(np.abs(rmin - omin) / omin) <= RADIUS_TOL_REL and np.max(np.abs(rmax - omax) / omax) <= RADIUS_TOL_REL
        off, _, _ = oracle.radius_search(pts, pts, r, max_nn=max_nn)
        assert k_gpu == int(off[-1])  # sum over queries of min(k, max_nn)


def test_pipeline_c1(ctx, oracle, c1):
    """normals -> RSD entirely on the GPU (fast mode) against the oracle pipeline.  Normals meet 1e-4 rad everywhere.
    The radii cannot meet 1e-4 relative everywhere END TO END in any arithmetic that is not the reference's bit for bit:
    its cosine is an fp32 expression (radius_estimation.cpp:153-155), and near |cos| = 1 one ulp of a normal moves the
    angle by up to 3.5e-4 rad.  The bound is therefore the oracle's own: the spread of its radii when its normals are
    moved by ONE ulp; the device path (normals within ~1e-6 rad of the oracle's) must stay well inside it."""
    r = 0.02
    ctx.upload(c1)
    ctx.build_grid(r)
    n4 = ctx.normals(r)
    rmin, rmax = ctx.rsd(r)
    o4, _ = oracle.normals(c1, r)
    omin, omax, _ = oracle.rsd(c1, o4, r)
    good = ~np.isnan(o4[:, 0])
    assert _angle(n4[good, :3], o4[good, :3]).max() <= NORMAL_TOL_RAD
    rel = np.maximum(np.abs(rmin - omin) / omin, np.abs(rmax - omax) / omax)
    rng = np.random.default_rng(7)
    bumped = o4.copy()
    sign = rng.integers(0, 2, size=(c1.shape[0], 3)).astype(np.float32) * 2 - 1
    bumped[:, :3] = np.nextafter(o4[:, :3], o4[:, :3] + sign)
    bmin, bmax, _ = oracle.rsd(c1, bumped, r)
    self_rel = np.maximum(np.abs(bmin - omin) / omin, np.abs(bmax - omax) / omax)
    print(f"pipeline: device vs oracle: {(rel > 1e-4).sum()} points over 1e-4 (max {rel.max():.3e}); "
          f"oracle vs oracle with normals 1 ulp off: {(self_rel > 1e-4).sum()} (max {self_rel.max():.3e})")
    # well inside the reference's own conditioning margin (measured: 0.14 % of the points against 1.6 %)
    assert np.mean(rel > RADIUS_TOL_REL) <= 0.5 * np.mean(self_rel > RADIUS_TOL_REL)
    assert np.median(rel) < 1e-6
    # and with identical normals the radii agree everywhere
    ctx.set_normals(o4)
    smin, smax = ctx.rsd(r)
    assert np.max(np.abs(smin - omin) / omin) <= RADIUS_TOL_REL and np.max(np.abs(smax - omax) / omax) <= RADIUS_TOL_REL


def test_pipeline_exact_mode(ctx_exact, oracle):
    pts = synth.tabletop(30_000, noise_sigma=0.0003)
    r = 0.02
    ctx_exact.upload(pts)
    ctx_exact.build_grid(r)
    n4 = ctx_exact.normals(r)
    rmin, rmax = ctx_exact.rsd(r)
    o4, _ = oracle.normals(pts, r)
    omin, omax, _ = oracle.rsd(pts, o4, r)
    assert np.mean(np.all(n4 == o4, axis=1)) > 0.98
    assert np.max(np.abs(rmin - omin) / omin) < RADIUS_TOL_REL
    assert np.max(np.abs(rmax - omax) / omax) < RADIUS_TOL_REL


def test_edge_cases(ctx, oracle):
    # empty cloud
    ctx.upload(np.zeros((0, 3), np.float32))
    ctx.build_grid(0.02)
    assert ctx.normals(0.02).shape == (0, 4)
    rmin, rmax = ctx.rsd(0.02)
    assert rmin.size == 0
    # single point, duplicates, NaN / inf points
    pts = np.array([[0, 0, 0], [0, 0, 0], [0.001, 0, 0], [0, 0.001, 0], [np.nan, 0, 0], [0.5, 0.5, 0.5],
                    [np.inf, 0, 0], [0.001, 0.001, 0.0005]], np.float32)
    ctx.upload(pts)
    ctx.build_grid(0.02)
    n4 = ctx.normals(0.02)
    o4, ok = oracle.normals(pts, 0.02)
    assert np.array_equal(np.isnan(n4[:, 0]), np.isnan(o4[:, 0]))
    off, idx, d2 = ctx.neighbors(0.02, 0, pts.shape[0])
    ooff, oidx, od2 = oracle.radius_search(pts, pts, 0.02)
    assert np.array_equal(off, ooff)
    gi, gd = _canon(off, idx, d2)
    assert np.array_equal(gi, oidx)
    ctx.set_normals(np.nan_to_num(o4[:, :3], nan=0.5))
    rmin, rmax = ctx.rsd(0.02)
    omin, omax, _ = oracle.rsd(pts, np.nan_to_num(o4[:, :3], nan=0.5), 0.02)
    assert np.allclose(rmin, omin, rtol=1e-4) and np.allclose(rmax, omax, rtol=1e-4)
    # radius larger than the grid cell is rejected, RSD without normals is rejected
    with pytest.raises(cab.CabError):
        ctx.normals(0.05)
    ctx.upload(pts)
    ctx.build_grid(0.02)
    with pytest.raises(cab.CabError, match="missing normals"):
        ctx.rsd(0.02)


def test_grsd_batch_bit_exact(ctx_exact, oracle):
    xyz, off = synth.clusters(28, 1300, 6000)
    leaf = 0.025
    hist = ctx_exact.grsd_batch(xyz, off, leaf, r_normals=0.02)
    vox = ctx_exact.grsd_voxels(len(off) - 1)
    mism = 0
    for c in range(len(off) - 1):
        o = oracle.grsd21(xyz[off[c]:off[c + 1]], leaf, r_normals=0.02)
        v0, v1 = vox["offsets"][c], vox["offsets"][c + 1]
        assert v1 - v0 == o["nvox"]
        og = oracle.voxel_grid(xyz[off[c]:off[c + 1]], leaf)
        assert np.array_equal(vox["centroids"][v0:v1].view(np.uint32), og["centroids"].view(np.uint32))
        assert np.allclose(vox["r_min"][v0:v1], o["radii"][:, 0], rtol=1e-5)
        assert np.allclose(vox["r_max"][v0:v1], o["radii"][:, 1], rtol=1e-5)
        if not np.array_equal(vox["labels"][v0:v1], o["labels"]):
            mism += 1
            continue
        assert np.array_equal(hist[c], o["hist21"])
    assert mism == 0


def test_grsd_batch_given_normals_and_ragged(ctx, oracle):
    xyz, off = synth.clusters(6, 1300, 3000, seed_extra=1)
    # add an empty cluster and a 1-point cluster
    xyz = np.concatenate([xyz, np.array([[0.3, 0.3, 0.9]], np.float32)])
    off = np.concatenate([off[:3], [off[3]], off[3:], [xyz.shape[0]]]).astype(np.int32)
    nc = len(off) - 1
    nrm = np.zeros((xyz.shape[0], 3), np.float32)
    for c in range(nc):
        if off[c + 1] > off[c]:
            nrm[off[c]:off[c + 1]] = oracle.normals(xyz[off[c]:off[c + 1]], 0.02)[0][:, :3]
    nrm = np.nan_to_num(nrm, nan=0.0)
    hist = ctx.grsd_batch(xyz, off, 0.025, normals=nrm)
    for c in range(nc):
        if off[c + 1] == off[c]:
            assert not hist[c].any()
            continue
        o = oracle.grsd21(xyz[off[c]:off[c + 1]], 0.025, normals_in=nrm[off[c]:off[c + 1]])
        assert np.array_equal(hist[c], o["hist21"]), c


def test_properties_at_full_size(ctx):
    """C2-sized run: size-independent properties instead of an oracle comparison."""
    pts = synth.scan(1_000_000)
    r = 0.03
    ctx.upload(pts)
    ctx.build_grid(r)
    n4 = ctx.normals(r)
    p1 = ctx.profile()
    rmin, rmax = ctx.rsd(r)
    p2 = ctx.profile()
    assert p1["neighbour_sum"] == p2["neighbour_sum"]  # both passes see the same neighbour sets
    assert p1["n_valid"] == pts.shape[0]
    good = ~np.isnan(n4[:, 0])
    assert np.allclose(np.linalg.norm(n4[good, :3], axis=1), 1.0, atol=1e-5)
    assert np.all(np.sum(n4[good, :3] * (-pts[good]), axis=1) >= -1e-4)  # oriented towards the origin
    assert np.all(rmin <= np.float32(0.1)) and np.all(rmax <= np.float32(0.1)) and np.all(rmin > 0)
    # idempotence and permutation invariance (lattice input: sums are exact only in exact mode,
    # so compare within tolerance)
    perm = np.random.default_rng(3).permutation(pts.shape[0])
    ctx.upload(pts[perm])
    ctx.build_grid(r)
    n4b = ctx.normals(r)
    rminb, rmaxb = ctx.rsd(r)
    assert ctx.profile()["neighbour_sum"] == p1["neighbour_sum"]
    ang = _angle(n4[perm][good[perm], :3], n4b[good[perm], :3])
    assert np.percentile(ang, 99.9) < 1e-4
    assert np.mean(np.abs(rminb - rmin[perm]) / rmin[perm] > 1e-4) < 2e-3


def test_sharded_run_equals_unsharded(ctx):
    """Query sharding (cab_set_shard): every shard's own slice equals the single-GPU result bit for bit;
    the shards are run one after the other on this GPU.  Normals of the rows around a shard are
    recomputed locally (halo), so no exchange between the passes is needed."""
    pts = synth.tabletop(60_000, noise_sigma=0.0002)
    r = 0.02
    ctx.set_shard(0, 1)
    ctx.upload(pts)
    ctx.build_grid(r)
    n4 = ctx.normals(r)
    rmin, rmax = ctx.rsd(r)
    seen = np.zeros(pts.shape[0], bool)
    W = 3
    try:
        for g in range(W):
            ctx.set_shard(g, W)
            ctx.upload(pts)
            ctx.build_grid(r)
            ctx.normals(r, download=False)
            ctx.rsd(r, download=False)
            b, e = ctx.shard_range()
            s4, srr, idx = ctx.download_sorted(b, e)
            assert not seen[idx].any()
            seen[idx] = True
            assert np.array_equal(s4.view(np.uint32), n4[idx].view(np.uint32))
            assert np.array_equal(srr[:, 0], rmin[idx]) and np.array_equal(srr[:, 1], rmax[idx])
    finally:
        ctx.set_shard(0, 1)
    assert seen.all()


def test_fused_normals_rsd_equals_two_calls(ctx):
    """cab_normals_rsd (normals copied out on the second stream while RSD runs) returns the bits of
    cab_normals followed by cab_rsd, in input order and as a shard's sorted slice."""
    pts = synth.tabletop(50_000, noise_sigma=0.0003)
    pts[17] = np.nan  # an invalid point keeps its index
    r = 0.02
    ctx.set_shard(0, 1)
    ctx.upload(pts)
    ctx.build_grid(r)
    n4 = ctx.normals(r)
    rmin, rmax = ctx.rsd(r)
    f4, fmin, fmax = ctx.normals_rsd(r)
    assert np.array_equal(f4.view(np.uint32), n4.view(np.uint32))
    assert np.array_equal(fmin.view(np.uint32), rmin.view(np.uint32))
    assert np.array_equal(fmax.view(np.uint32), rmax.view(np.uint32))
    # max_nn on both passes
    n4t = ctx.normals(r, max_nn=40)
    rmint, rmaxt = ctx.rsd(r, max_nn=75)
    t4, tmin, tmax = ctx.normals_rsd(r, max_nn_normals=40, max_nn_rsd=75)
    assert np.array_equal(t4.view(np.uint32), n4t.view(np.uint32))
    assert np.array_equal(tmin, rmint) and np.array_equal(tmax, rmaxt)
    try:
        seen = np.zeros(pts.shape[0], bool)
        for g in range(2):
            ctx.set_shard(g, 2)
            ctx.upload(pts)
            ctx.build_grid(r)
            s4, srr, idx = ctx.normals_rsd(r, sorted_shard=True)
            seen[idx] = True
            assert np.array_equal(s4.view(np.uint32), n4[idx].view(np.uint32))
            assert np.array_equal(srr[:, 0], rmin[idx]) and np.array_equal(srr[:, 1], rmax[idx])
        assert seen.sum() == pts.shape[0] - 1  # the NaN point belongs to no shard slice
    finally:
        ctx.set_shard(0, 1)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_grsd_signature_variants_bit_exact(ctx_exact, oracle, kind):
    """GRSD-21 with subdivisions, GRSD-325 and PlusGRSD-110 (cab_grsd_signatures) against the oracle
    recipe, per cluster, bit-exact; includes a cluster smaller than the offsets and an empty one."""
    xyz, off = synth.clusters(10, 1300, 5000, seed_extra=2)
    off = np.concatenate([off[:4], [off[4]], off[4:]]).astype(np.int32)  # an empty cluster
    nc = len(off) - 1
    leaf = 0.025
    # both sides get the same point normals (a few of them NaN: such voxels count as "to empty")
    nrm = np.zeros((xyz.shape[0], 3), np.float32)
    for c in range(nc):
        if off[c + 1] > off[c]:
            nrm[off[c]:off[c + 1]] = oracle.normals(xyz[off[c]:off[c + 1]], 0.02)[0][:, :3]
    nrm[np.random.default_rng(kind).integers(0, xyz.shape[0], 40)] = np.nan
    hist21 = ctx_exact.grsd_batch(xyz, off, leaf, normals=nrm)
    vox = ctx_exact.grsd_voxels(nc)
    for sub, o in [(0, (0, 0, 0)), (2, (0, 0, 0)), (2, (1, 0, 1)), (3, (4, 0, 0))]:
        got = ctx_exact.grsd_signatures(nc, kind, sub, o)
        assert got["hist"].shape[1] == cab.SIG_DIM[kind]
        for c in range(nc):
            h0, h1 = got["offsets"][c], got["offsets"][c + 1]
            pts = xyz[off[c]:off[c + 1]]
            if pts.shape[0] == 0:
                assert not got["hist"][h0:h1].any()
                continue
            want = oracle.grsd_cluster(pts, leaf, kind, sub, o, normals_in=nrm[off[c]:off[c + 1]])
            v0, v1 = vox["offsets"][c], vox["offsets"][c + 1]
            assert np.array_equal(vox["labels"][v0:v1], want["labels"])
            assert h1 - h0 == max(want["hist_num"], 0), (c, sub, o)
            if want["hist_num"] > 0:
                assert np.array_equal(got["subdiv_b"][c], want["subdiv_b"])
                assert np.array_equal(got["hist"][h0:h1], want["hist"]), (c, sub, o)
        if kind == 0 and sub == 0:
            assert np.array_equal(got["hist"], hist21)
    # the sliding boxes partition the voxels when the offsets are zero
    whole = ctx_exact.grsd_signatures(nc, kind, 0)["hist"]
    boxes = ctx_exact.grsd_signatures(nc, kind, 2)
    summed = np.stack([boxes["hist"][boxes["offsets"][c]:boxes["offsets"][c + 1]].sum(0) for c in range(nc)])
    assert np.array_equal(summed, whole)
    # state is tied to the cloud: a new upload invalidates it
    ctx_exact.upload(xyz[:100])
    with pytest.raises(cab.CabError, match="cab_grsd_batch first"):
        ctx_exact.grsd_signatures(nc, kind)


def test_fast_normals_count_neighbours_exactly_at_the_boundary(ctx, oracle):
    """The fast normals pass decides most candidates with the centred-monomial test and re-tests the
    ones near the radius with the exact rule: neighbour counts must still equal the oracle's on a
    cloud with thousands of pairs exactly at, and one ulp beyond, the radius."""
    r = np.float32(0.0625)
    rng = np.random.default_rng(9)
    base = synth.quantize(rng.uniform(0.2, 0.6, size=(6000, 3)))
    extra = []
    for p in base[:1500]:
        extra.append(p + np.array([r, 0, 0], np.float32))
        extra.append(p + np.array([0, np.nextafter(r, np.float32(1)), 0], np.float32))
        extra.append(p - np.array([0, 0, r], np.float32))
        extra.append(p + np.array([0, -np.nextafter(r, np.float32(0)), 0], np.float32))
    pts = np.concatenate([base, np.array(extra, np.float32)]).astype(np.float32)
    for cloud in (pts, pts + np.float32(37.5), synth.tabletop(30_000, noise_sigma=0.0005)):
        rr = float(r) if cloud is not None and cloud.shape[0] == pts.shape[0] else 0.02
        ctx.upload(cloud)
        ctx.build_grid(rr)
        n4 = ctx.normals(rr)
        o4, ok = oracle.normals(cloud, rr)
        assert ctx.profile()["neighbour_sum"] == int(ok.sum())
        good = ~np.isnan(o4[:, 0])
        assert np.array_equal(np.isnan(n4[:, 0]), ~good)
        assert np.mean(_angle(n4[good, :3], o4[good, :3]) > 1e-3) < 1e-2  # random 3D cloud: ill-conditioned normals allowed


def test_properties_at_c4_size(ctx):
    """The benchmark cloud itself (C4: 20 M points, r = 2 cm): size-independent properties, since the oracle would need
    most of a minute for it.  Determinism, the two passes seeing the same neighbour sets, unit normals oriented to the
    viewpoint, radii inside (0, plane_radius], the fused call returning the same bits, and an oracle check on a slab."""
    pts = synth.room(20_000_000)
    r = 0.02
    ctx.set_shard(0, 1)
    ctx.upload(pts)
    ctx.build_grid(r)
    n4 = ctx.normals(r)
    p1 = ctx.profile()
    rmin, rmax = ctx.rsd(r)
    p2 = ctx.profile()
    assert p1["n_valid"] == pts.shape[0] and p1["neighbour_sum"] == p2["neighbour_sum"]
    assert 200 < p1["neighbour_sum"] / pts.shape[0] < 300  # the density the workload is specified at (k ~ 250)
    good = ~np.isnan(n4[:, 0])
    assert good.mean() > 0.9999
    assert np.allclose(np.linalg.norm(n4[good, :3], axis=1), 1.0, atol=1e-5)
    assert np.all(np.sum(n4[good, :3] * (-pts[good]), axis=1) >= -1e-4)
    assert np.all(rmin > 0) and np.all(rmin <= np.float32(0.1)) and np.all(rmax <= np.float32(0.1))
    # the same bits from a second, fused run
    f4, fmin, fmax = ctx.normals_rsd(r)
    assert np.array_equal(f4.view(np.uint32), n4.view(np.uint32))
    assert np.array_equal(fmin, rmin) and np.array_equal(fmax, rmax)
    # neighbour sets of a few queries in the middle of the cloud, bit-exact against brute force in numpy
    q0 = 10_000_000
    off, idx, d2 = ctx.neighbors(r, q0, q0 + 16)
    for t in range(16):
        d = pts - pts[q0 + t]
        dd = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]  # fp32, the documented rule
        want = np.flatnonzero(dd <= np.float32(r) * np.float32(r))
        got = np.sort(idx[off[t]:off[t + 1]])
        assert np.array_equal(got, want)


@pytest.mark.parametrize("k_mean", [16, 512])
def test_density_patches_neighbour_sets(ctx, oracle, k_mean):
    """The two ends of the C5 density sweep (mean 16 and 512 neighbours per query): neighbour sets of a slab of queries
    bit-exact against the oracle."""
    pts = synth.density_patches(300_000, float(k_mean), 0.02)
    r = 0.02
    ctx.set_shard(0, 1)
    ctx.upload(pts)
    ctx.build_grid(r)
    q0, q1 = 150_000, 152_000
    off, idx, d2 = ctx.neighbors(r, q0, q1)
    ooff, oidx, od2 = oracle.radius_search(pts, pts[q0:q1], r)
    assert np.array_equal(off, ooff)
    gi, gd = _canon(off, idx, d2)
    assert np.array_equal(gi, oidx) and np.array_equal(gd.view(np.uint32), od2.view(np.uint32))
    assert 0.7 * k_mean < (off[-1] / (q1 - q0)) < 1.3 * k_mean
    # the sparse end gets cells that are wide along x (the table would otherwise hold tens of cells per point); the
    # dense end keeps cubic cells
    ext = pts.max(axis=0).astype(np.float64) - pts.min(axis=0).astype(np.float64)
    cubic = np.prod(np.floor(ext / (r * (1 + 1 / 1024))) + 2)
    n_cells = ctx.profile()["n_cells"]
    if k_mean == 16:
        assert cubic > 16 * pts.shape[0] and n_cells < cubic / 4, (cubic, n_cells)
    else:
        assert n_cells > cubic / 2, (cubic, n_cells)


def test_wide_cells_normals_and_radii(oracle):
    """A sparse cloud (mean 16 neighbours): with the cell table 16 edges wide along x, normals and radii still are what the
    oracle says for every point."""
    pts = synth.density_patches(300_000, 16.0, 0.02)
    r = 0.02
    c = cab.Context(0)
    c.upload(pts)
    c.build_grid(r)
    ext = pts.max(axis=0).astype(np.float64) - pts.min(axis=0).astype(np.float64)
    assert c.profile()["n_cells"] < np.prod(np.floor(ext / (r * (1 + 1 / 1024))) + 2) / 4
    n4 = c.normals(r)
    o4, ok = oracle.normals(pts, r)
    assert c.profile()["neighbour_sum"] == int(ok.sum())
    good = ~np.isnan(o4[:, 0])
    assert np.array_equal(np.isnan(n4[:, 0]), ~good)
    ang = np.linalg.norm(np.cross(n4[good, :3].astype(np.float64), o4[good, :3].astype(np.float64)), axis=1)
    assert np.mean(ang > NORMAL_TOL_RAD) < 5e-3  # (sparse neighbourhoods: a few near-degenerate covariances)
    c.set_normals(np.nan_to_num(o4[:, :3], nan=0.0))
    rmin, rmax = c.rsd(r)
    omin, omax, _ = oracle.rsd(pts, np.nan_to_num(o4[:, :3], nan=0.0), r)
    assert np.max(np.abs(rmin - omin) / omin) <= RADIUS_TOL_REL and np.max(np.abs(rmax - omax) / omax) <= RADIUS_TOL_REL
    c.close()


def test_spread_out_cloud_gets_coarser_cells(oracle):
    """Two small objects 300 m apart with a 1 cm radius: a dense table of 1 cm cells would need ~10^13 entries; the grid is
    built with coarser cells and the neighbour sets, normals and radii stay what the oracle says."""
    a = synth.tabletop(20_000, noise_sigma=0.0003)
    b = synth.tabletop(20_000, noise_sigma=0.0003, seed_extra=1) + np.array([300.0, -150.0, 40.0], np.float32)
    pts = np.concatenate([a, b]).astype(np.float32)
    r = 0.01
    c = cab.Context(0, max_table_cells=1 << 22)
    c.upload(pts)
    c.build_grid(r)
    assert c.profile()["n_cells"] <= 1 << 22
    off, idx, d2 = c.neighbors(r, 0, pts.shape[0])
    ooff, oidx, od2 = oracle.radius_search(pts, pts, r)
    assert np.array_equal(off, ooff)
    gi, gd = _canon(off, idx, d2)
    assert np.array_equal(gi, oidx) and np.array_equal(gd.view(np.uint32), od2.view(np.uint32))
    n4 = c.normals(r)
    o4, ok = oracle.normals(pts, r)
    assert c.profile()["neighbour_sum"] == int(ok.sum())
    c.set_normals(np.nan_to_num(o4[:, :3], nan=0.0))
    rmin, rmax = c.rsd(r)
    omin, omax, _ = oracle.rsd(pts, np.nan_to_num(o4[:, :3], nan=0.0), r)
    assert np.max(np.abs(rmin - omin) / omin) <= RADIUS_TOL_REL and np.max(np.abs(rmax - omax) / omax) <= RADIUS_TOL_REL
    c.close()


@pytest.mark.parametrize("n", [5, 1001, 1002, 1003, 4096])
def test_grid_build_layouts_agree(ctx, oracle, n):
    """The vector key / bounds passes (packed xyz, 16-byte aligned, 4 points per thread) and the scalar passes
    (stride 4, a device pointer that is only 4-byte aligned) build the same grid: neighbour sets bit-exact for
    cloud sizes with every remainder mod 4, including non-finite points in the tail."""
    import torch

    full = synth.tabletop(20_000, noise_sigma=0.0004)
    near = np.argsort(np.linalg.norm(full - full[123], axis=1), kind="stable")[:n]  # a dense patch of n points
    pts = np.ascontiguousarray(full[np.sort(near)])
    if n > 8:
        pts[n - 1] = np.nan
        pts[n - 2, 1] = np.inf
    r = 0.02
    ooff, oidx, _ = oracle.radius_search(pts, pts, r)

    def sets():
        off, idx, d2 = ctx.neighbors(r, 0, n)
        return off, _canon(off, idx, d2)[0]

    ctx.upload(pts)  # vector layout
    ctx.build_grid(r)
    off, gi = sets()
    assert np.array_equal(off, ooff) and np.array_equal(gi, oidx)
    p4 = np.zeros((n, 4), np.float32)
    p4[:, :3] = pts
    p4[:, 3] = 7.0
    ctx.upload(p4)  # stride 4: scalar layout
    ctx.build_grid(r)
    off, gi = sets()
    assert np.array_equal(off, ooff) and np.array_equal(gi, oidx)
    dev = torch.zeros(3 * n + 1, dtype=torch.float32, device="cuda:0")
    dev[1:] = torch.from_numpy(pts.reshape(-1)).to("cuda:0")
    torch.cuda.synchronize()
    ctx.set_cloud_device(dev.data_ptr() + 4, n, 3)  # packed but misaligned: scalar layout
    ctx.build_grid(r)
    off, gi = sets()
    assert np.array_equal(off, ooff) and np.array_equal(gi, oidx)
    for g in range(2):  # and as one shard of two (sharded sort over the vector passes)
        ctx.set_shard(g, 2)
        try:
            ctx.upload(pts)
            ctx.build_grid(r)
            b, e = ctx.shard_range()
            ctx.normals(r, download=False)
            s4, _, sidx = ctx.download_sorted(b, e, rsd=False)
        finally:
            ctx.set_shard(0, 1)
        o4, _ = oracle.normals(pts, r)
        assert np.array_equal(np.isnan(s4[:, 0]), np.isnan(o4[sidx, 0]))


@pytest.mark.parametrize("exact", [False, True])
def test_normals_against_the_references_own_output(ctx, ctx_exact, exact):
    """The device normals against the normals and curvatures the reference itself wrote into
    color_chlac/demos/data/tmp_normal.pcd (computeNormal, radius 0.02, viewpoint 0; tests/golden/tmp_normal.npz)."""
    import pathlib

    g = np.load(pathlib.Path(__file__).resolve().parent / "golden" / "tmp_normal.npz")
    xyz, ref_n, ref_c = g["xyz"], g["normal"], g["curvature"]
    c = ctx_exact if exact else ctx
    c.set_shard(0, 1)
    c.upload(xyz)
    c.build_grid(0.02)
    n4 = c.normals(0.02)
    assert (np.sum(n4[:, :3].astype(np.float64) * ref_n, axis=1) > 0).all()  # same side of the surface everywhere
    ang = _angle(n4[:, :3], ref_n)
    assert np.sum(ang > 1e-4) <= 2 and np.percentile(ang, 99.9) < 5e-5
    assert np.abs(n4[:, 3] - ref_c).max() < (1e-6 if exact else 2e-5)
